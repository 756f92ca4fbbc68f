"""World-size-2 `gloo` test (CPU) of the slice-parallel host logic: the contiguous slice blocks
of `qtn_contract_sliced` partition the slice set, every rank derives the identical plan, and the
per-rank partial sums (oracle arithmetic here -- there is no GPU) all-reduce to the full amplitude."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as graft
    from oracle import contract as oc, network as on, plan as oplan
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q = graft.load_package()
    net, _, _ = q.circuits.cfg2_network(10, 8, seed=5)
    q.optimize_contraction_order(net)                       # C++ order search on every rank
    il = q.contract_rep(net)
    arrays = [t.data for t in net.tensors]
    shapes = [a.shape for a in arrays]
    S = q.choose_slices(shapes, il, None, 5, 8)              # C++ slice chooser on every rank
    plan = q.ContractionPlan(shapes, il, None, S)
    n = plan.nslices
    # every rank must hold the same plan: compare a digest through the process group
    digest = torch.tensor([float(sum(S)), float(n), plan.flops_per_slice, float(plan.nsteps)], dtype=torch.float64)
    lo, hi = digest.clone(), digest.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi)
    s0, s1 = n * rank // world, n * (rank + 1) // world       # the rule of qtn_contract_sliced_range
    part = complex(oplan.contract_sliced(arrays, il, None, S, range(s0, s1)))
    buf = torch.tensor([part.real, part.imag], dtype=torch.float64)
    dist.all_reduce(buf)                                       # the one exchange step of the path
    cover = torch.zeros(n, dtype=torch.int64)
    cover[s0:s1] = 1
    dist.all_reduce(cover)
    full = complex(oc.contract(on.Network([on.Tensor(a) for a in arrays], [on.Summation(s.idx) for s in net.contractions], [])))
    ok = bool(torch.all(cover == 1)) and abs(complex(buf[0], buf[1]) - full) < 1e-12 * abs(full)
    # EXTENSION: the searched order (qtn_order_search) is deterministic, so every rank derives the same tree and
    # slice set without any exchange, and the slice-parallel sum over that tree gives the same amplitude
    net2, _, _ = q.circuits.cfg2_network(10, 8, seed=5)
    il2 = q.contract_rep(net2)
    order, info = q.search_order(shapes, il2, 32, 3, 5)
    S2 = q.choose_slices(shapes, il2, order, 5, 1)
    n2 = 2 ** len(S2)
    digest2 = torch.tensor([float(sum((i + 1) * l for i, l in enumerate(order))), float(sum(S2)), info["total_flops"]], dtype=torch.float64)
    lo, hi = digest2.clone(), digest2.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    t0, t1 = n2 * rank // world, n2 * (rank + 1) // world
    part2 = complex(oplan.contract_sliced(arrays, il2, order, S2, range(t0, t1)))
    buf2 = torch.tensor([part2.real, part2.imag], dtype=torch.float64)
    dist.all_reduce(buf2)
    ok = ok and torch.equal(lo, hi) and n2 > 1 and abs(complex(buf2[0], buf2[1]) - full) < 1e-12 * abs(full)
    ret[rank] = ok
    dist.destroy_process_group()


def test_slice_parallel_world2_gloo():
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)
