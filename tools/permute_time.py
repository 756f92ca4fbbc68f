"""K1 permute kernel: achieved HBM GB/s (read + write bytes / CUDA-event time) on the GPU box."""
import ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
import torch
q = graft.load_package()
from qaintensor_b200 import _lib
_lib.require_device()
ext = torch.cuda.ExternalStream(_lib.stream_ptr())
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6521.4
rng = np.random.default_rng(0)
cases = [("28 dim-2 modes, random perm", (2,) * 28, list(rng.permutation(28) + 1)),
         ("28 dim-2 modes, reverse", (2,) * 28, list(range(28, 0, -1))),
         ("26 dim-2 modes, swap halves", (2,) * 26, list(range(14, 27)) + list(range(1, 14))),
         ("matrix transpose 16384 x 16384", (16384, 16384), [2, 1]),
         ("switch! (chi,2,2,chi) chi=4096", (4096, 2, 2, 4096), [1, 3, 2, 4]),
         ("contract_svd (l,2,r)->(2,r,l) 4096", (4096, 2, 4096), [2, 3, 1]),
         ("12 modes of 4..8", (8, 4, 6, 4, 8, 5, 4, 7, 4, 6, 4, 4), [7, 3, 12, 1, 9, 5, 2, 11, 4, 8, 10, 6])]
rows = []
for name, shape, perm in cases:
    n = int(np.prod(shape))
    a = torch.randn(2 * n, dtype=torch.float64, device="cuda")
    b = torch.empty_like(a)
    dims, pm = _lib.arr_i64(shape), _lib.arr_i32(perm)
    for _ in range(3):
        _lib.check(_lib.lib.qtn_permutedims_device(a.data_ptr(), len(shape), dims, pm, 0, b.data_ptr()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record(ext)
    for _ in range(reps):
        _lib.check(_lib.lib.qtn_permutedims_device(a.data_ptr(), len(shape), dims, pm, 0, b.data_ptr()))
    e1.record(ext)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gbs = 2 * n * 16 / (ms * 1e-3) / 1e9
    # correctness spot check against torch
    # correctness: random element spot check (column-major index arithmetic)
    oshape = [shape[p - 1] for p in perm]
    idx = [rng.integers(0, d, size=64) for d in oshape]
    ostr = np.cumprod([1] + list(oshape[:-1])); istr = np.cumprod([1] + list(shape[:-1]))
    oo = sum(i * s for i, s in zip(idx, ostr)); ii = sum(idx[k] * istr[perm[k] - 1] for k in range(len(perm)))
    av = a.view(-1, 2)[torch.as_tensor(ii, device="cuda")]; bv = b.view(-1, 2)[torch.as_tensor(oo, device="cuda")]
    ok = torch.equal(av, bv)
    rows.append(dict(case=name, elems=n, ms=ms, gbs=gbs, frac=gbs / peak, ok=bool(ok)))
    print("%-40s %10d elems  %8.3f ms  %7.0f GB/s  %.2f of %.0f  ok=%s" % (name, n, ms, gbs, gbs / peak, peak, ok))
json.dump(rows, open("gpurun_out/permute_time.json", "w"), indent=1)
