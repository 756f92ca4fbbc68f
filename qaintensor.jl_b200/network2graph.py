"""``network2graph`` surface (src/network2graph.jl): the order search runs in the
library's host C++ (``qtn_order_treewidth``), bit-exact with the reference."""
import ctypes as C
import warnings

from ._lib import arr_i32, check, lib


class Graph:
    """Minimal simple graph (sorted adjacency lists, 1-based) for the exported
    ``network_graph`` / ``line_graph`` (src/network2graph.jl:53-73, 121-150)."""

    def __init__(self, n=0):
        self.adj = [[] for _ in range(n)]

    def nv(self):
        return len(self.adj)

    def ne(self):
        loops = sum(1 for v, l in enumerate(self.adj, 1) if v in l)
        return (sum(len(l) for l in self.adj) + loops) // 2

    def add_vertex(self):
        self.adj.append([])

    def add_edge(self, s, d):
        if d in self.adj[s - 1]:
            return False
        self.adj[s - 1].append(d)
        self.adj[s - 1].sort()
        if s != d:
            self.adj[d - 1].append(s)
            self.adj[d - 1].sort()
        return True

    def neighbors(self, v):
        return self.adj[v - 1]

    def degree(self):
        return [len(l) for l in self.adj]

    def edges(self):
        return [(s, d) for s in range(1, self.nv() + 1) for d in self.adj[s - 1] if d >= s]


def network_graph(net):
    G = Graph(len(net.tensors))
    edge_idx = {}
    for k, s in enumerate(net.contractions, 1):
        if len(s.idx) != 2:
            raise ValueError("Contractions of more than 2 tensors not supported")
        i, j = sorted((s.idx[0][0], s.idx[1][0]))
        G.add_edge(i, j)
        edge_idx.setdefault((i, j), []).append(k)
    return G, edge_idx


def line_graph(net):
    if len(net.openidx) != 0:
        warnings.warn("All open indices are disregarded")
    G, edge_idx = network_graph(net)
    LG = Graph()
    nodeinfo = []
    for i in range(1, G.nv() + 1):
        for j in G.neighbors(i):
            if j > i:
                for e in edge_idx[(i, j)]:
                    LG.add_vertex()
                    nodeinfo.append((i, j, e))
    for a in range(len(nodeinfo)):
        for b in range(a + 1, len(nodeinfo)):
            if set(nodeinfo[a][:2]) & set(nodeinfo[b][:2]):
                LG.add_edge(a + 1, b + 1)
    return LG, nodeinfo


def _pairs(net):
    flat = []
    for s in net.contractions:
        if len(s.idx) != 2:
            raise ValueError("Contractions of more than 2 tensors not supported")
        flat += [s.idx[0][0], s.idx[0][1], s.idx[1][0], s.idx[1][1]]
    return flat


def contraction_order_perm(net):
    """Permutation ``perm`` with ``net.contractions[perm]`` = the treewidth order
    (src/network2graph.jl:429-446), and the width of the line-graph decomposition."""
    nc = len(net.contractions)
    perm = (C.c_int32 * max(nc, 1))()
    tw = C.c_int32(0)
    check(lib.qtn_order_treewidth(len(net.tensors), nc, arr_i32(_pairs(net)), perm, C.byref(tw)))
    return [int(perm[i]) for i in range(nc)], int(tw.value)


def contraction_order(net):
    perm, _ = contraction_order_perm(net)
    out = []
    for k in perm:
        i, j = sorted((net.contractions[k - 1].idx[0][0], net.contractions[k - 1].idx[1][0]))
        out.append((i, j, k))
    return out


def optimize_contraction_order(net, method="treewidth", ntrials=256, seed=0, max_log2_elems=-1):
    """``optimize_contraction_order!(net)`` (src/network2graph.jl:473-479).

    EXTENSION: ``method="search"`` replaces the reference's treewidth heuristic by the
    randomised-greedy + annealing search of ``qtn_order_search`` (same contract: ``net.contractions`` is
    permuted in place so that the default ascending-label walk of ``contract`` follows the found tree;
    open indices are handled).  The default stays the reference's order."""
    if method == "search":
        from .contract import contract_rep, search_order
        order, _ = search_order([t.data.shape for t in net.tensors], contract_rep(net), ntrials, seed, max_log2_elems)
        net.contractions = [net.contractions[k - 1] for k in order]
        return None
    if method != "treewidth":
        raise ValueError("method must be 'treewidth' (reference) or 'search' (extension)")
    if len(net.openidx) != 0:
        warnings.warn("For TensorNetworks with open indices the treewidth algorithm is unlikely to optimize performance")
        warnings.warn("All open indices are disregarded")
    perm, _ = contraction_order_perm(net)
    net.contractions = [net.contractions[k - 1] for k in perm]
    return None


def tree_decomposition_width(nv, edges):
    """Width and min-fill elimination order of the reference's heuristic
    (src/network2graph.jl:224-272, 300-337) for a plain graph given by 1-based edges."""
    flat = [x for e in edges for x in e]
    tw = C.c_int32(0)
    order = (C.c_int32 * max(nv, 1))()
    check(lib.qtn_graph_treewidth(nv, len(edges), arr_i32(flat), C.byref(tw), order))
    return int(tw.value), [int(order[i]) for i in range(nv)]
