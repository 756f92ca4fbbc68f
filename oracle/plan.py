"""Oracle: pairwise contraction tree, cost model, slicing rule (test infrastructure).

The tree is the symbolic form of the ``ncon`` walk restated in ``contract.py``
(``src/contract.jl:257, 263`` -> TensorOperations 3.1.0).  Slicing is an
EXTENSION with no reference counterpart; the rule below
is the definition both the oracle and the product's C++ planner implement,
and the product must reproduce the chosen label set bit-exactly.

Slice rule (deterministic greedy):
  repeat while  max tensor size > 2**max_log2_elems  or  #slices < min_slices:
    candidates = un-sliced contracted labels (extent > 1) present on at least
                 one tensor of maximal size (inputs or intermediates of the tree);
    choose the candidate minimising (total_flops_after, max_size_after, label);
  slice id s in [0, prod extents): label S[j] takes digit j of s, S[0] fastest.
Every slice runs the *same* tree with the sliced labels dropped from all label
lists (a pair left without shared labels becomes an outer product, K = 1).
"""
import itertools

import numpy as np


def contraction_tree(network, order=None):
    """Symbolic ncon walk.  Returns (nodes, steps).

    nodes[i] = label list of node i (inputs first, traces already removed);
    steps = [(a, b, out, shared_labels)], node ``out`` has labels free(a)+free(b).
    """
    nodes = []
    for lab in network:
        lab = list(lab)
        nodes.append([l for l in lab if lab.count(l) == 1])
    flat = [l for lab in nodes for l in lab]
    if order is None:
        order = sorted(set(l for l in flat if l > 0))
    alive = list(range(len(nodes)))
    steps = []

    def merge(a, b):
        la, lb = nodes[a], nodes[b]
        shared = [l for l in la if l in lb]
        out = [l for l in la if l not in shared] + [l for l in lb if l not in shared]
        nodes.append(out)
        o = len(nodes) - 1
        steps.append((a, b, o, shared))
        alive[alive.index(a)] = o
        alive.remove(b)

    for lab in order:
        holders = [n for n in alive if lab in nodes[n]]
        if len(holders) == 2:
            merge(holders[0], holders[1])
    while len(alive) > 1:
        merge(alive[0], alive[1])
    return nodes, steps


def label_dims(tensors, network):
    dims = {}
    for t, lab in zip(tensors, network):
        for d, l in zip(np.shape(t), lab):
            dims[l] = int(d)
    return dims


def tree_cost(nodes, steps, dims, sliced=()):
    """(flops, bytes, max_elems, per-step (M, N, K)) for ONE slice, sliced labels dropped.

    flops = sum 8*M*N*K, bytes = sum 16*(M*K + K*N + M*N)  (SURVEY.md section 8d).
    """
    sl = set(sliced)

    def size(labs):
        s = 1
        for l in labs:
            if l not in sl:
                s *= dims[l]
        return s

    flops = 0
    nbytes = 0
    mx = max([size(n) for n in nodes[:len(nodes) - len(steps)]] + [1])
    mnk = []
    for a, b, o, shared in steps:
        K = size(shared)
        M = size(nodes[a]) // K
        N = size(nodes[b]) // K
        flops += 8 * M * N * K
        nbytes += 16 * (M * K + K * N + M * N)
        mx = max(mx, M * N)
        mnk.append((M, N, K))
    return flops, nbytes, mx, mnk


def choose_slice_labels(nodes, steps, dims, max_log2_elems=28, min_slices=1):
    sliced = []
    limit = 2 ** max_log2_elems

    def nslices(S):
        n = 1
        for l in S:
            n *= dims[l]
        return n

    def size(labs, sl):
        s = 1
        for l in labs:
            if l not in sl:
                s *= dims[l]
        return s

    while True:
        _, _, mx, _ = tree_cost(nodes, steps, dims, sliced)
        if mx <= limit and nslices(sliced) >= min_slices:
            break
        sl = set(sliced)
        cand = set()
        for labs in nodes:
            if size(labs, sl) == mx:
                cand.update(l for l in labs if l > 0 and l not in sl and dims[l] > 1)
        if not cand:
            break
        best = None
        for l in sorted(cand):
            f, _, m2, _ = tree_cost(nodes, steps, dims, sliced + [l])
            key = (f * nslices(sliced + [l]), m2, l)
            if best is None or key < best:
                best = key
        sliced.append(best[2])
    return sliced


def _index_fixed(arr, labs, fixed):
    idx = tuple(fixed[l] if l in fixed else slice(None) for l in labs)
    return arr[idx], [l for l in labs if l not in fixed]


def execute_tree(tensors, network, nodes, steps, fixed=None, stats=None):
    """Run the tree numerically; ``fixed`` = {label: 0-based index value} for sliced labels."""
    from .contract import _trace_repeated, pairwise_ttgt
    fixed = fixed or {}
    vals = []
    for arr, lab in zip(tensors, network):
        # fix the sliced labels first: a sliced label that sits twice on one tensor (self-contraction) then reads the
        # diagonal element, and the sum over the slices is the trace; only the remaining repeats are traced here
        a, l = _index_fixed(np.asarray(arr), list(lab), fixed)
        vals.append(_trace_repeated(a, l))
    vals += [None] * len(steps)
    for a, b, o, _ in steps:
        A, la = vals[a]
        B, lb = vals[b]
        vals[o] = pairwise_ttgt(A, la, B, lb, stats)
        vals[a] = vals[b] = None
    arr, lab = vals[-1] if steps else vals[0]
    out_order = sorted(lab, reverse=True)
    return np.transpose(arr, [lab.index(l) for l in out_order]) if lab else arr


def slice_assignment(slice_labels, dims, sid):
    fixed = {}
    for l in slice_labels:
        fixed[l] = sid % dims[l]
        sid //= dims[l]
    return fixed


def contract_sliced(tensors, network, order=None, slice_labels=(), slice_ids=None):
    """Sum over slices (all, or the given ids) of the per-slice tree contraction."""
    nodes, steps = contraction_tree(network, order)
    dims = label_dims(tensors, network)
    n = 1
    for l in slice_labels:
        n *= dims[l]
    total = None
    for sid in (range(n) if slice_ids is None else slice_ids):
        part = execute_tree(tensors, network, nodes, steps, slice_assignment(slice_labels, dims, sid))
        total = part if total is None else total + part
    return total
