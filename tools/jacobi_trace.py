"""Analyses a CTA trace of the Jacobi round kernels (QTN_JACOBI_TRACE=<file>, one sweep):
    python tools/jacobi_trace.py trace.bin
Per kernel kind: CTA count and duration statistics; per SM: busy fractions; over time (50 us bins): the number of
resident CTAs of each kind -- i.e. how well the sub-batch streams overlap the eigensolves with the tensor-pipe kernels."""
import sys

import numpy as np

KINDS = ["gram_full", "gram_cross", "eig", "update"]
rec = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 3)
kind = (rec[:, 0] & 0xFF).astype(int)
smid = ((rec[:, 0] >> 8) & 0xFFFF).astype(int)
t0 = rec[:, 1].astype(np.int64)
t1 = rec[:, 2].astype(np.int64)
base = t0.min()
t0 -= base
t1 -= base
span = t1.max()
print("records %d, span %.3f ms, SMs seen %d" % (len(rec), span / 1e6, len(np.unique(smid))))
for k, name in enumerate(KINDS):
    m = kind == k
    if not m.any():
        continue
    d = (t1[m] - t0[m]) / 1e3
    print("%-10s n=%7d  dur us: mean %.1f  p10 %.1f  p50 %.1f  p90 %.1f  max %.1f   CTA-time total %.1f ms (= %.2f CTAs resident on average per SM)"
          % (name, m.sum(), d.mean(), np.percentile(d, 10), np.percentile(d, 50), np.percentile(d, 90), d.max(), d.sum() / 1e3,
             d.sum() * 1e3 / span / 148))
# residency over time
nb = int(span // 50000) + 1
edges = np.arange(nb + 1) * 50000
print("\ntime-resolved average resident CTAs per SM (50 us bins, first 40 bins):")
print("  bin_us  " + "  ".join("%10s" % n for n in KINDS))
occ = np.zeros((4, nb))
for k in range(4):
    m = kind == k
    for a, b in zip(t0[m], t1[m]):
        i0, i1 = int(a // 50000), int(b // 50000)
        if i0 == i1:
            occ[k, i0] += b - a
        else:
            occ[k, i0] += edges[i0 + 1] - a
            occ[k, i0 + 1:i1] += 50000
            occ[k, i1] += b - edges[i1]
occ /= 50000.0 * 148
for i in range(min(nb, 40)):
    print("  %6d  " % (i * 50) + "  ".join("%10.2f" % occ[k, i] for k in range(4)))
print("\nmean over the sweep: " + ", ".join("%s %.2f" % (KINDS[k], occ[k].mean()) for k in range(4)))
# per-SM idle time: union of intervals of DMMA kernels (kinds 0, 1, 3)
idle = []
for s in np.unique(smid):
    m = (smid == s) & (kind != 2)
    iv = sorted(zip(t0[m], t1[m]))
    busy, cur_a, cur_b = 0, None, None
    for a, b in iv:
        if cur_b is None or a > cur_b:
            if cur_b is not None:
                busy += cur_b - cur_a
            cur_a, cur_b = a, b
        else:
            cur_b = max(cur_b, b)
    if cur_b is not None:
        busy += cur_b - cur_a
    idle.append(1 - busy / span)
print("fraction of the sweep during which an SM holds NO tensor-pipe CTA: mean %.3f, min %.3f, max %.3f" % (np.mean(idle), np.min(idle), np.max(idle)))
