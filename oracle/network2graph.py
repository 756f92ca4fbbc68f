"""Oracle restatement of ``src/network2graph.jl`` (test infrastructure).

Integer / graph work only; the result must be reproduced bit-exactly by the
product's C++ host code (``qtn_order_treewidth``).  Each function cites the
reference lines it restates; LightGraphs semantics live in ``lightgraphs.py``.
"""
import warnings

from .lightgraphs import Graph, induced_subgraph, is_connected


def subset(A, B):  # src/network2graph.jl:4-9  (the reference's `⊂`)
    try:
        it = iter(A)
    except TypeError:
        it = iter((A,))  # Julia numbers iterate as one element
    for a in it:
        if a not in B:
            return False
    return True


def local_circuit_graph(N, k):  # src/network2graph.jl:35-43
    G = Graph(N)
    for i in range(1, N):
        for j in range(1, min(k - 1, N - i) + 1):
            G.add_edge(i, i + j)
    return G


def random_graph(Nn, Ne, rng):  # src/network2graph.jl:16-28 (sampling via numpy rng)
    if not Ne <= Nn * (Nn - 1) // 2:
        raise ValueError("Number of edges must be smaller or equal than N(N-1)/2, with N the number of vertices")
    G = Graph(Nn)
    possible = [(i, j) for i in range(1, Nn + 1) for j in range(1, i)]
    for t in rng.choice(len(possible), size=Ne, replace=False):
        G.add_edge(*possible[int(t)])
    return G


def network_graph(net):  # src/network2graph.jl:53-73
    M = len(net.tensors)
    G = Graph(M)
    edge_idx = {}
    for k, s in enumerate(net.contractions, 1):
        if len(s.idx) != 2:
            raise ValueError("Contractions of more than 2 tensors not supported")
        i, j = s.idx[0][0], s.idx[1][0]
        if not i <= j:
            i, j = j, i
        G.add_edge(i, j)
        edge_idx.setdefault((i, j), []).append(k)
    return G, edge_idx


def line_graph_of_graph(G):  # src/network2graph.jl:80-107
    LG = Graph()
    nodeinfo = []
    for i in range(1, G.nv() + 1):
        for j in G.neighbors(i):
            edge = (i, j) if i < j else (j, i)
            if edge not in nodeinfo:
                LG.add_vertex()
                nodeinfo.append(edge)
    for i, e1 in enumerate(nodeinfo, 1):
        for j, e2 in enumerate(nodeinfo[i:], 1):
            if set(e1) & set(e2):
                LG.add_edge(i, j + i)
    return LG, nodeinfo


def line_graph(net):  # src/network2graph.jl:121-150
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
    # add edges: nodes sharing a tensor (reference loop is O(E^2), :141-148;
    # same edge set built per shared tensor here -- add_edge keeps lists sorted)
    by_tensor = {}
    for n, (a, b, _) in enumerate(nodeinfo, 1):
        by_tensor.setdefault(a, []).append(n)
        by_tensor.setdefault(b, []).append(n)
    for nodes in by_tensor.values():
        for x in range(len(nodes)):
            for y in range(x + 1, len(nodes)):
                LG.add_edge(nodes[x], nodes[y])
    return LG, nodeinfo


def lacking_for_clique_neigh(G, i):  # src/network2graph.jl:182-191
    neigh = G.neighbors(i)
    lacking = []
    for j, i2 in enumerate(neigh):
        for i1 in neigh[:j]:
            if not G.has_edge(i1, i2):
                lacking.append((i1, i2))
    return len(lacking), lacking


def rem_vertex_fill(G, i, lacking, ordering, vertex_label):  # src/network2graph.jl:204-215
    for e in lacking:
        G.add_edge(e[0], e[1])
    ordering.append(vertex_label[i - 1])
    G.rem_vertex(i)
    v = vertex_label.pop()
    if i <= G.nv():
        vertex_label[i - 1] = v


def min_fill_ordering(G):  # src/network2graph.jl:224-272
    H = G.copy()
    ordering = []
    vertex_label = list(range(1, H.nv() + 1))
    while H.nv() > 0:
        success = False
        for i in range(H.nv(), 0, -1):  # range fixed at entry, degree queried live
            if H.degree(i) == 0:
                rem_vertex_fill(H, i, [], ordering, vertex_label)
                success = True
        for i in range(H.nv(), 0, -1):
            if H.degree(i) == 1:
                rem_vertex_fill(H, i, [], ordering, vertex_label)
                success = True
        if not success:
            degrees = H.degree()
            J = sorted(range(1, H.nv() + 1), key=lambda v: degrees[v - 1])  # stable sortperm
            found_clique = False
            v = 0
            best_n_lacking = float("inf")
            best_lacking = []
            for j in J:
                n_lacking, lacking = lacking_for_clique_neigh(H, j)
                if n_lacking == 0:
                    rem_vertex_fill(H, j, lacking, ordering, vertex_label)
                    found_clique = True
                    break
                elif n_lacking < best_n_lacking:
                    v = j
                    best_n_lacking = n_lacking
                    best_lacking = lacking
            if not found_clique:
                rem_vertex_fill(H, v, best_lacking, ordering, vertex_label)
    return ordering


def triangulation(G, ordering):  # src/network2graph.jl:280-292
    H = G.copy()
    pos = {v: i for i, v in enumerate(ordering)}
    for i, v in enumerate(ordering):
        high_neigh = [w for w in H.neighbors(v) if pos[w] > i]
        for j, i1 in enumerate(high_neigh):
            for i2 in high_neigh[:j]:
                H.add_edge(i1, i2)
    return H


def tree_decomposition(G):  # src/network2graph.jl:300-337
    ordering = min_fill_ordering(G)
    H = triangulation(G, ordering)
    n = H.nv()
    pos = {v: i for i, v in enumerate(ordering)}
    up_neighs = [[w for w in H.neighbors(ordering[i]) if pos[w] > i] for i in range(n)]
    # findfirst(length.(up_neighs) .== nv(H)-1:-1:0)
    clique_neigh_idx = next(i for i in range(n) if len(up_neighs[i]) == n - 1 - i)
    first_bag = list(up_neighs[clique_neigh_idx])
    if ordering[clique_neigh_idx] not in first_bag:
        first_bag.append(ordering[clique_neigh_idx])
    decomp = Graph(1)
    bags = [first_bag]
    tw = len(first_bag) - 1
    for i in range(clique_neigh_idx - 1, -1, -1):
        neigh = up_neighs[i]
        old_bag_idx = 0
        for j, bag in enumerate(bags, 1):
            if subset(neigh, bag):
                old_bag_idx = j
                break
        if old_bag_idx == 0:
            old_bag_idx = 1
        decomp.add_vertex()
        new_bag = list(neigh) + [ordering[i]]
        bags.append(new_bag)
        tw = max(tw, len(new_bag) - 1)
        decomp.add_edge(old_bag_idx, decomp.nv())
    return tw, decomp, bags


def is_tree_decomposition(G, tree, bags):  # src/network2graph.jl:344-382
    if sorted(set().union(*[set(b) for b in bags])) != list(range(1, G.nv() + 1)):
        warnings.warn("Union of bags is not equal to union of vertices")
        return False
    for e in G.edges():
        if not any(subset(e, B) for B in bags):
            warnings.warn("Edge (%d, %d) not found in any bag" % e)
            return False
    subgraphs = [[] for _ in range(G.nv())]
    for i, b in enumerate(bags, 1):
        for j in b:
            subgraphs[j - 1].append(i)
    for v, s in enumerate(subgraphs, 1):
        if len(s) > 0:
            subtree, _ = induced_subgraph(tree, s)
            if not is_connected(subtree):
                warnings.warn("Subgraph for vertex %d not connected" % v)
                return False
    return True


def contraction_order_graph(H, edges):  # src/network2graph.jl:391-422
    if len(edges) != H.nv():
        raise ValueError("Invalid list of edges for `H`; the length of `edges` must equal the number of vertices of `H`")
    tw, tree, bags = tree_decomposition(H)
    contr_order = []
    degrees = tree.degree()
    while max(degrees) > 0:
        leaves_idx = [i for i, d in enumerate(degrees, 1) if d == 1]
        lens = [len(bags[i - 1]) for i in leaves_idx]
        leaf_idx = leaves_idx[lens.index(min(lens))]  # argmin = first minimum
        leaf_bag = bags[leaf_idx - 1]
        neigh_idx = tree.neighbors(leaf_idx)[0]
        neigh_bag = bags[neigh_idx - 1]
        seen = set()
        for i in leaf_bag:  # setdiff keeps first-argument order, unique
            if i not in neigh_bag and i not in seen:
                seen.add(i)
                contr_order.append(edges[i - 1])
        tree.rem_vertex(leaf_idx)
        moved_bag = bags.pop()
        if leaf_idx <= tree.nv():
            bags[leaf_idx - 1] = moved_bag
        degrees = tree.degree()
    assert len(bags) == 1
    for i in bags[0]:
        contr_order.append(edges[i - 1])
    return contr_order


def contraction_order(net):  # src/network2graph.jl:429-446
    _, edge_idx = network_graph(net)
    H, edges = line_graph(net)
    con_order = contraction_order_graph(H, edges)
    auto_con = []
    for i in range(1, len(net.tensors) + 1):
        if (i, i) in edge_idx:
            for k in edge_idx[(i, i)]:
                auto_con.append((i, i, k))
    return auto_con + con_order


def optimize_contraction_order(net):  # src/network2graph.jl:473-479  (`optimize_contraction_order!`)
    if len(net.openidx) != 0:
        warnings.warn("For TensorNetworks with open indices the treewidth algorithm is unlikely to optimize performance")
    new_order = contraction_order(net)
    perm = [t[2] for t in new_order]
    net.contractions = [net.contractions[k - 1] for k in perm]
    return perm
