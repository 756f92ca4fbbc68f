"""Oracle restatement of ``src/contract.jl`` (test infrastructure).

``contract`` forwards to ``TensorOperations.ncon`` (v3.1.0, un-vendored; call
sites ``src/contract.jl:257, 263``).  ``ncon`` is restated here from the
published package semantics: labels are processed in
order; the two current sub-results holding a label are contracted pairwise
over *all* labels they share (TTGT: permute, reshape, ``zgemm``); labels already
internal are skipped; a label repeated on one input tensor is a partial trace
done first; leftover pieces are combined by outer product; output axes follow
the negative labels -1, -2, ...
"""
import numpy as np


# ----------------------------------------------------------------------------
# contract_rep  (src/contract.jl:8-32 and :39-60)
# ----------------------------------------------------------------------------
def contract_rep(net, optimize=False):
    indexlist = [[0] * t.ndims for t in net.tensors]
    leg_costs = {}
    for i, ts in enumerate(net.contractions, 1):
        for (tj, lj) in ts.idx:
            assert 1 <= lj <= net.tensors[tj - 1].ndims
            indexlist[tj - 1][lj - 1] = i
            leg_costs[i] = net.tensors[tj - 1].size[lj - 1]
    j = len(net.contractions)
    for i, (ot, ol) in enumerate(net.openidx, 1):
        assert indexlist[ot - 1][ol - 1] == 0
        indexlist[ot - 1][ol - 1] = -i - j
        leg_costs[i + j] = net.tensors[ot - 1].size[ol - 1]
    for idx in indexlist:
        for i in idx:
            assert i != 0
    if optimize:
        return leg_costs, indexlist
    return indexlist


# ----------------------------------------------------------------------------
# ncon (TensorOperations 3.1.0 semantics)
# ----------------------------------------------------------------------------
def _trace_repeated(arr, labels):
    """Partial trace over labels that occur twice on one tensor."""
    labels = list(labels)
    while True:
        dup = next((l for l in labels if labels.count(l) == 2), None)
        if dup is None:
            return arr, labels
        a = labels.index(dup)
        b = labels.index(dup, a + 1)
        arr = np.trace(arr, axis1=a, axis2=b)
        labels = [l for l in labels if l != dup]


def pairwise_ttgt(A, la, B, lb, stats=None):
    """One TTGT step: contract A and B over all shared labels.

    Returns (C, labels) with C's axes = (free of A in order, free of B in order).
    """
    shared = [l for l in la if l in lb]
    fa = [l for l in la if l not in shared]
    fb = [l for l in lb if l not in shared]
    pa = [la.index(l) for l in fa] + [la.index(l) for l in shared]
    pb = [lb.index(l) for l in shared] + [lb.index(l) for l in fb]
    da = [A.shape[la.index(l)] for l in fa]
    db = [B.shape[lb.index(l)] for l in fb]
    dk = [A.shape[la.index(l)] for l in shared]
    M = int(np.prod(da, dtype=np.int64)) if da else 1
    N = int(np.prod(db, dtype=np.int64)) if db else 1
    K = int(np.prod(dk, dtype=np.int64)) if dk else 1
    Am = np.transpose(A, pa).reshape(M, K)
    Bm = np.transpose(B, pb).reshape(K, N)
    C = (Am @ Bm).reshape(da + db)
    if stats is not None:
        stats.append((M, N, K))
    return C, fa + fb


def ncon(tensors, network, order=None, stats=None):
    """``TensorOperations.ncon(tensors, network; order)``.

    ``tensors``: numpy arrays; ``network[i][j]``: label of leg j of tensor i
    (positive = contracted, appears twice; negative = open, appears once).
    ``stats`` (optional list) receives one (M, N, K) per pairwise GEMM.
    """
    groups = []
    for arr, lab in zip(tensors, network):
        arr = np.asarray(arr)
        assert arr.ndim == len(lab)
        a, l = _trace_repeated(arr, lab)
        groups.append([a, l])
    flat = [l for _, lab in groups for l in lab]
    for l in set(flat):
        assert l != 0
        assert flat.count(l) == (2 if l > 0 else 1), "not a valid ncon network"
    if order is None:
        order = sorted(set(l for l in flat if l > 0))
    for lab in order:
        holders = [g for g in groups if lab in g[1]]
        if len(holders) < 2:
            continue  # already contracted as a shared label of an earlier step
        g1, g2 = holders
        C, lc = pairwise_ttgt(g1[0], g1[1], g2[0], g2[1], stats)
        i1 = next(i for i, g in enumerate(groups) if g is g1)
        groups[i1] = [C, lc]
        groups = [g for g in groups if g is not g2]
    # disconnected pieces: outer products
    while len(groups) > 1:
        g1, g2 = groups[0], groups[1]
        C, lc = pairwise_ttgt(g1[0], g1[1], g2[0], g2[1], stats)
        groups = [[C, lc]] + groups[2:]
    arr, lab = groups[0]
    out_order = sorted(lab, reverse=True)  # -1, -2, ...
    return np.transpose(arr, [lab.index(l) for l in out_order]) if lab else arr


# ----------------------------------------------------------------------------
# contract_order: exhaustive cost-capped search  (src/contract.jl:68-235)
# ----------------------------------------------------------------------------
def _build_cost(freelegs, commonlegs, leg_costs, mu_old, mu, isnew, cost1, cost2):
    # src/contract.jl:68-86
    new_cost = 1
    for i in range(len(freelegs)):
        if freelegs[i] or commonlegs[i]:
            new_cost *= leg_costs[i + 1]
    new_cost += cost1 + cost2
    if new_cost > mu:
        return new_cost, False
    if not isnew:
        if new_cost <= mu_old:
            return float("inf"), False
    return new_cost, True


def contract_order(net, leg_costs, indexlist):
    """src/contract.jl:184-235 with check_contraction! (:96-160) inlined."""
    n = len(net.tensors)
    numlegs = len(leg_costs)
    S = [[] for _ in range(n)]          # S[c-1]: objects built from c tensors
    newflags = [[] for _ in range(n)]
    for i, x in enumerate(indexlist):
        leg = [False] * numlegs
        for l in x:
            leg[abs(l) - 1] = True
        tf = [False] * n
        tf[i] = True
        S[0].append(dict(leg=leg, tf=tf, seq=[], cost=0))
    newflags[0] = [True] * n
    mu_old, mu_cap, mu_next = 0, 1, float("inf")
    while len(S[n - 1]) == 0:
        for c in range(2, n + 1):
            for d in range(1, c // 2 + 1):
                a, b = d, c - d
                Sa, Sb, Sab = S[a - 1], S[b - 1], S[a + b - 1]
                for i, Ta in enumerate(Sa):
                    for j, Tb in enumerate(Sb):
                        if any(x and y for x, y in zip(Ta["tf"], Tb["tf"])):
                            continue
                        common = [x and y for x, y in zip(Ta["leg"], Tb["leg"])]
                        free = [x != y for x, y in zip(Ta["leg"], Tb["leg"])]
                        if not any(common):
                            continue
                        new_cost, ok = _build_cost(free, common, leg_costs, mu_old, mu_cap,
                                                   newflags[a - 1][i] or newflags[b - 1][j],
                                                   Ta["cost"], Tb["cost"])
                        if not ok:
                            mu_next = min(mu_next, new_cost)
                            continue
                        tin = [x or y for x, y in zip(Ta["tf"], Tb["tf"])]
                        objptr = next((p for p, o in enumerate(Sab) if o["tf"] == tin), None)
                        seq = Ta["seq"] + Tb["seq"] + [p + 1 for p, cflag in enumerate(common) if cflag]
                        obj = dict(leg=free, tf=tin, seq=seq, cost=new_cost)
                        if objptr is None:
                            Sab.append(obj)
                            newflags[a + b - 1].append(True)
                        elif Sab[objptr]["cost"] > new_cost:
                            Sab[objptr] = obj
                            newflags[a + b - 1][objptr] = True
        mu_old, mu_cap, mu_next = mu_cap, mu_next, float("inf")
        for f in newflags:
            for p in range(len(f)):
                f[p] = False
    return S[n - 1][0]["seq"], S[n - 1][0]["cost"]


# ----------------------------------------------------------------------------
# contract  (src/contract.jl:242-264)
# ----------------------------------------------------------------------------
def contract(net, optimize=False, stats=None):
    if len(net.tensors) == 1:
        return np.transpose(net.tensors[0].data, [l - 1 for (_, l) in net.openidx])
    if optimize:
        leg_costs, indexlist = contract_rep(net, True)
        sequence, _ = contract_order(net, leg_costs, indexlist)
        # quirk Q1 (:250-257): labels are renamed to their position in `sequence`
        # AND `order=sequence` is passed on top of the renamed labels.
        for lab in indexlist:
            for j in range(len(lab)):
                if lab[j] > 0:
                    lab[j] = sequence.index(lab[j]) + 1
        return ncon([t.data for t in net.tensors], indexlist, order=sequence, stats=stats)
    indexlist = contract_rep(net)
    return ncon([t.data for t in net.tensors], indexlist, stats=stats)
