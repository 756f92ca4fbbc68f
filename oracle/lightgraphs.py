"""Minimal emulation of LightGraphs 1.3.5 ``SimpleGraph`` (test infrastructure).

LightGraphs is an un-vendored dependency of the reference (``Manifest.toml``
pins 1.3.5).  The contraction order the reference emits depends on a handful of
its behaviours, restated here from the published package:

* adjacency lists are kept **sorted ascending**; ``neighbors(g, v)`` returns
  that list; ``degree(g, v)`` is its length (a self-loop counts once);
* ``add_edge!`` returns ``false`` for an existing edge / out-of-range vertex;
* ``rem_vertex!(g, v)`` removes v's edges and then **moves the last vertex n
  into slot v** (its edges re-attached, lists re-sorted) before popping.

Vertices are 1-based to match the reference's call sites
(``src/network2graph.jl`` throughout).
"""
from bisect import bisect_left


class Graph:
    def __init__(self, n=0):
        self.adj = [[] for _ in range(n)]  # adj[v-1] sorted list of 1-based ids

    # --- queries -----------------------------------------------------
    def nv(self):
        return len(self.adj)

    def ne(self):
        loops = sum(1 for v, l in enumerate(self.adj, 1) if v in l)
        return (sum(len(l) for l in self.adj) + loops) // 2

    def neighbors(self, v):
        return self.adj[v - 1]

    def degree(self, v=None):
        if v is None:
            return [len(l) for l in self.adj]
        return len(self.adj[v - 1])

    def has_edge(self, s, d):
        if not (1 <= s <= self.nv() and 1 <= d <= self.nv()):
            return False
        l = self.adj[s - 1]
        i = bisect_left(l, d)
        return i < len(l) and l[i] == d

    def edges(self):
        return [(s, d) for s in range(1, self.nv() + 1) for d in self.adj[s - 1] if d >= s]

    def copy(self):
        g = Graph()
        g.adj = [list(l) for l in self.adj]
        return g

    def __eq__(self, other):
        return isinstance(other, Graph) and self.adj == other.adj

    # --- mutation ----------------------------------------------------
    def add_vertex(self):
        self.adj.append([])
        return True

    def add_edge(self, s, d):
        n = self.nv()
        if not (1 <= s <= n and 1 <= d <= n):
            return False
        l = self.adj[s - 1]
        i = bisect_left(l, d)
        if i < len(l) and l[i] == d:
            return False
        l.insert(i, d)
        if s == d:
            return True
        l = self.adj[d - 1]
        l.insert(bisect_left(l, s), s)
        return True

    def rem_edge(self, s, d):
        if not self.has_edge(s, d):
            return False
        self.adj[s - 1].remove(d)
        if s != d:
            self.adj[d - 1].remove(s)
        return True

    def rem_vertex(self, v):
        n = self.nv()
        if not (1 <= v <= n):
            return False
        for s in list(self.adj[v - 1]):
            self.rem_edge(s, v)
        neigs = list(self.adj[n - 1])
        for s in neigs:
            self.rem_edge(s, n)
        self_loop_n = False
        if v != n:
            for s in neigs:
                if s != n:
                    self.add_edge(s, v)
                else:
                    self_loop_n = True
        if self_loop_n:
            self.add_edge(v, v)
        self.adj.pop()
        return True


def complete_graph(n):
    g = Graph(n)
    for i in range(1, n + 1):
        for j in range(i + 1, n + 1):
            g.add_edge(i, j)
    return g


def is_connected(g):
    n = g.nv()
    if n == 0:
        return True
    seen = {1}
    stack = [1]
    while stack:
        v = stack.pop()
        for w in g.neighbors(v):
            if w not in seen:
                seen.add(w)
                stack.append(w)
    return len(seen) == n


def induced_subgraph(g, vlist):
    """LightGraphs ``induced_subgraph(g, vlist)``: vertex i of the result is vlist[i]."""
    pos = {v: i + 1 for i, v in enumerate(vlist)}
    h = Graph(len(vlist))
    for v in vlist:
        for w in g.neighbors(v):
            if w in pos:
                h.add_edge(pos[v], pos[w])
    return h, list(vlist)
