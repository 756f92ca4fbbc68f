// Host-side contraction-order search (integer-only, bit-exact with the reference).
//
// Restates src/network2graph.jl:53-73 (network_graph), :121-150 (line_graph),
// :182-215 (lacking_for_clique_neigh, rem_vertex_fill!), :224-272
// (min_fill_ordering), :280-292 (triangulation), :300-337 (tree_decomposition),
// :391-446 (contraction_order) and src/contract.jl:68-235 (contract_order).
// The LightGraphs 1.3.5 SimpleGraph behaviours the result depends on (sorted
// adjacency, swap-with-last rem_vertex!, stable sortperm, first argmin) are
// implemented by SGraph below.  Vertices are 0-based internally.
#include "qtn_internal.h"

#include <algorithm>
#include <map>
#include <numeric>

namespace qtn {

struct SGraph {
    std::vector<std::vector<int>> adj;  // sorted ascending
    // optional bit-matrix mirror of adj (enable_bits): row v = W words; kept current by add_edge / rem_edge, so the
    // swap-with-last relabelling of rem_vertex (built from those two) carries over.  Used by min_fill_ordering only.
    std::vector<uint64_t> bits;
    int W = 0;
    explicit SGraph(int n = 0) : adj(n) {}
    int nv() const { return (int)adj.size(); }
    int degree(int v) const { return (int)adj[v].size(); }
    bool has_edge(int s, int d) const { return std::binary_search(adj[s].begin(), adj[s].end(), d); }
    void enable_bits() {
        W = (nv() + 63) / 64;
        bits.assign((size_t)nv() * W, 0);
        for (int v = 0; v < nv(); ++v)
            for (int w : adj[v]) bits[(size_t)v * W + (w >> 6)] |= (uint64_t)1 << (w & 63);
    }
    const uint64_t* row(int v) const { return bits.data() + (size_t)v * W; }
    void set_bit(int s, int d, bool on) {
        uint64_t& w = bits[(size_t)s * W + (d >> 6)];
        const uint64_t m = (uint64_t)1 << (d & 63);
        w = on ? (w | m) : (w & ~m);
    }
    bool add_edge(int s, int d) {
        if (s < 0 || d < 0 || s >= nv() || d >= nv()) return false;
        auto& ls = adj[s];
        auto it = std::lower_bound(ls.begin(), ls.end(), d);
        if (it != ls.end() && *it == d) return false;
        ls.insert(it, d);
        if (W) set_bit(s, d, true);
        if (s == d) return true;
        auto& ld = adj[d];
        ld.insert(std::lower_bound(ld.begin(), ld.end(), s), s);
        if (W) set_bit(d, s, true);
        return true;
    }
    void rem_edge(int s, int d) {
        auto& ls = adj[s];
        auto it = std::lower_bound(ls.begin(), ls.end(), d);
        if (it == ls.end() || *it != d) return;
        ls.erase(it);
        if (W) set_bit(s, d, false);
        if (s != d) {
            auto& ld = adj[d];
            ld.erase(std::lower_bound(ld.begin(), ld.end(), s));
            if (W) set_bit(d, s, false);
        }
    }
    // LightGraphs rem_vertex!: drop v's edges, move the last vertex into slot v.
    void rem_vertex(int v) {
        int n = nv() - 1;
        std::vector<int> srcs = adj[v];
        for (int s : srcs) rem_edge(s, v);
        std::vector<int> neigs = adj[n];
        for (int s : neigs) rem_edge(s, n);
        bool self_loop_n = false;
        if (v != n) {
            for (int s : neigs) {
                if (s != n) add_edge(s, v);
                else self_loop_n = true;
            }
        }
        if (self_loop_n) add_edge(v, v);
        adj.pop_back();
    }
};

static void lacking_for_clique_neigh(const SGraph& G, int i, std::vector<std::pair<int, int>>& lacking) {
    lacking.clear();
    const auto& neigh = G.adj[i];
    for (size_t j = 0; j < neigh.size(); ++j)
        for (size_t a = 0; a < j; ++a)
            if (!G.has_edge(neigh[a], neigh[j])) lacking.emplace_back(neigh[a], neigh[j]);
}

static void rem_vertex_fill(SGraph& G, int i, const std::vector<std::pair<int, int>>& lacking,
                            std::vector<int>& ordering, std::vector<int>& vertex_label) {
    for (auto& e : lacking) G.add_edge(e.first, e.second);
    ordering.push_back(vertex_label[i]);
    G.rem_vertex(i);
    int v = vertex_label.back();
    vertex_label.pop_back();
    if (i < G.nv()) vertex_label[i] = v;
}

// Number of missing edges among the neighbours of i (= lacking_for_clique_neigh(G, i).size()).
static int lacking_count(const SGraph& G, int i) {
    const auto& neigh = G.adj[i];
    if (G.W) {   // d (d - 1) / 2 - #edges inside N(i), the latter from row intersections (no self-loops in this mode)
        const uint64_t* ri = G.row(i);
        long twice = 0;
        for (int a : neigh) {
            const uint64_t* ra = G.row(a);
            for (int w = 0; w < G.W; ++w) twice += __builtin_popcountll(ra[w] & ri[w]);
        }
        const long d = (long)neigh.size();
        return (int)(d * (d - 1) / 2 - twice / 2);
    }
    int c = 0;
    for (size_t j = 0; j < neigh.size(); ++j)
        for (size_t a = 0; a < j; ++a) c += !G.has_edge(neigh[a], neigh[j]);
    return c;
}

// min_fill_ordering (src/network2graph.jl:224-272).  The reference recomputes the lacking edges of EVERY vertex at
// every elimination (in sortperm(degree) order: the first vertex whose neighbourhood is a clique, else the first one
// with the fewest lacking edges).  The choices here are the same -- the scan takes the minimum in (degree, index)
// order, which is what the stable sortperm + first-minimum rule selects -- but the counts are cached per vertex:
// eliminating x changes the adjacency of N(x) only (recomputed), and a fill edge (a, b) takes exactly one lacking
// pair from every other common neighbour of a and b (decremented); rem_vertex!'s swap-with-last relabelling moves
// the cached entry with the vertex.  Line graph of cfg 3 (516 vertices, width 54): 290 -> ~25 ms.
static std::vector<int> min_fill_ordering(const SGraph& G) {
    SGraph H = G;
    bool loops = false;
    for (int v = 0; v < H.nv(); ++v) loops = loops || H.has_edge(v, v);
    if (!loops && H.nv() <= 16384) H.enable_bits();
    std::vector<int> ordering;
    std::vector<int> vertex_label(H.nv());
    std::iota(vertex_label.begin(), vertex_label.end(), 0);
    const std::vector<std::pair<int, int>> none;
    std::vector<std::pair<int, int>> lacking;
    std::vector<int> cnt(H.nv(), -1);   // cached lacking count per current vertex index; -1 = stale
    auto eliminate = [&](int i, const std::vector<std::pair<int, int>>& fill) {
        for (int a : H.adj[i]) cnt[a] = -1;
        for (auto& e : fill) {   // every common neighbour w of a new edge's ends loses the lacking pair (a, b)
            if (H.W) {
                const uint64_t *ra = H.row(e.first), *rb = H.row(e.second);
                for (int w = 0; w < H.W; ++w)
                    for (uint64_t m = ra[w] & rb[w]; m; m &= m - 1) {
                        const int x = w * 64 + __builtin_ctzll(m);
                        if (cnt[x] > 0) --cnt[x];
                    }
            } else {
                for (int x : H.adj[e.first]) if (H.has_edge(x, e.second) && cnt[x] > 0) --cnt[x];
            }
        }
        rem_vertex_fill(H, i, fill, ordering, vertex_label);
        const int moved = cnt.back();   // rem_vertex! moved the last vertex into slot i
        cnt.pop_back();
        if (i < H.nv()) cnt[i] = moved;
    };
    while (H.nv() > 0) {
        bool success = false;
        for (int i = H.nv() - 1; i >= 0; --i)
            if (H.degree(i) == 0) { eliminate(i, none); success = true; }
        for (int i = H.nv() - 1; i >= 0; --i)
            if (H.degree(i) == 1) { eliminate(i, none); success = true; }
        if (success) continue;
        const int n = H.nv();
        int clique = -1, v = -1;
        for (int j = 0; j < n; ++j) {
            if (cnt[j] < 0) cnt[j] = lacking_count(H, j);
            if (cnt[j] == 0) {
                if (clique < 0 || H.degree(j) < H.degree(clique)) clique = j;
            } else if (v < 0 || cnt[j] < cnt[v] || (cnt[j] == cnt[v] && H.degree(j) < H.degree(v))) {
                v = j;
            }
        }
        const int pick = clique >= 0 ? clique : v;
        lacking_for_clique_neigh(H, pick, lacking);
        eliminate(pick, lacking);
    }
    return ordering;
}

struct TreeDecomp {
    int tw;
    SGraph tree;
    std::vector<std::vector<int>> bags;
    std::vector<int> ordering;
};

static TreeDecomp tree_decomposition(const SGraph& G) {
    TreeDecomp td;
    td.ordering = min_fill_ordering(G);
    const auto& ordering = td.ordering;
    int n = G.nv();
    std::vector<int> pos(n);
    for (int i = 0; i < n; ++i) pos[ordering[i]] = i;
    // triangulation (src/network2graph.jl:280-292)
    SGraph H = G;
    for (int i = 0; i < n; ++i) {
        std::vector<int> high;
        for (int w : H.adj[ordering[i]]) if (pos[w] > i) high.push_back(w);
        for (size_t j = 0; j < high.size(); ++j)
            for (size_t a = 0; a < j; ++a) H.add_edge(high[j], high[a]);
    }
    std::vector<std::vector<int>> up(n);
    for (int i = 0; i < n; ++i)
        for (int w : H.adj[ordering[i]]) if (pos[w] > i) up[i].push_back(w);
    int cidx = -1;
    for (int i = 0; i < n; ++i) if ((int)up[i].size() == n - 1 - i) { cidx = i; break; }
    std::vector<int> first_bag = up[cidx];
    if (std::find(first_bag.begin(), first_bag.end(), ordering[cidx]) == first_bag.end())
        first_bag.push_back(ordering[cidx]);
    td.tree = SGraph(1);
    td.bags.push_back(first_bag);
    td.tw = (int)first_bag.size() - 1;
    std::vector<char> mark(n, 0);
    for (int i = cidx - 1; i >= 0; --i) {
        const auto& neigh = up[i];
        int old_bag = -1;
        for (size_t j = 0; j < td.bags.size(); ++j) {
            const auto& bag = td.bags[j];
            if (bag.size() < neigh.size()) continue;
            for (int x : bag) mark[x] = 1;
            bool sub = true;
            for (int x : neigh) if (!mark[x]) { sub = false; break; }
            for (int x : bag) mark[x] = 0;
            if (sub) { old_bag = (int)j; break; }
        }
        if (old_bag < 0) old_bag = 0;
        td.tree.adj.emplace_back();
        std::vector<int> nb = neigh;
        nb.push_back(ordering[i]);
        td.tw = std::max(td.tw, (int)nb.size() - 1);
        td.bags.push_back(std::move(nb));
        td.tree.add_edge(old_bag, td.tree.nv() - 1);
    }
    return td;
}

// contraction_order(H, edges) (src/network2graph.jl:391-422): returns line-graph
// vertices (0-based) in contraction order.
static std::vector<int> order_from_decomposition(TreeDecomp& td) {
    std::vector<int> out;
    SGraph& tree = td.tree;
    auto& bags = td.bags;
    std::vector<char> inb;
    while (true) {
        int maxdeg = 0;
        for (int v = 0; v < tree.nv(); ++v) maxdeg = std::max(maxdeg, tree.degree(v));
        if (maxdeg == 0) break;
        int leaf = -1;
        size_t best = (size_t)-1;
        for (int v = 0; v < tree.nv(); ++v)
            if (tree.degree(v) == 1 && bags[v].size() < best) { best = bags[v].size(); leaf = v; }
        int nb = tree.adj[leaf][0];
        std::map<int, char> innb;
        for (int x : bags[nb]) innb[x] = 1;
        std::map<int, char> seen;
        for (int x : bags[leaf])
            if (!innb.count(x) && !seen.count(x)) { seen[x] = 1; out.push_back(x); }
        tree.rem_vertex(leaf);
        std::vector<int> moved = std::move(bags.back());
        bags.pop_back();
        if (leaf < tree.nv()) bags[leaf] = std::move(moved);
    }
    for (int x : bags[0]) out.push_back(x);
    return out;
}

int order_treewidth(int ntensors, int ncontr, const int32_t* pairs, int32_t* perm_out, int32_t* tw_out) {
    if (ntensors <= 0 || ncontr < 0) return fail(QTN_EINVAL, "qtn_order_treewidth: bad sizes");
    // network_graph (src/network2graph.jl:53-73)
    SGraph G(ntensors);
    std::map<std::pair<int, int>, std::vector<int>> edge_idx;
    for (int k = 0; k < ncontr; ++k) {
        int i = pairs[4 * k] - 1, j = pairs[4 * k + 2] - 1;
        if (i < 0 || j < 0 || i >= ntensors || j >= ntensors)
            return fail(QTN_EINVAL, "qtn_order_treewidth: tensor index out of range");
        if (i > j) std::swap(i, j);
        G.add_edge(i, j);
        edge_idx[{i, j}].push_back(k);
    }
    // line_graph(net) (src/network2graph.jl:121-150)
    std::vector<std::array<int, 3>> nodeinfo;
    for (int i = 0; i < ntensors; ++i)
        for (int j : G.adj[i])
            if (j > i)
                for (int e : edge_idx[{i, j}]) nodeinfo.push_back({i, j, e});
    int nn = (int)nodeinfo.size();
    std::vector<int> perm;
    // self-contractions first (src/network2graph.jl:436-445)
    for (int i = 0; i < ntensors; ++i) {
        auto it = edge_idx.find({i, i});
        if (it != edge_idx.end()) for (int k : it->second) perm.push_back(k);
    }
    int tw = 0;
    if (nn > 0) {
        SGraph LG(nn);
        std::vector<std::vector<int>> by_tensor(ntensors);
        for (int n = 0; n < nn; ++n) { by_tensor[nodeinfo[n][0]].push_back(n); by_tensor[nodeinfo[n][1]].push_back(n); }
        for (auto& nodes : by_tensor)
            for (size_t x = 0; x < nodes.size(); ++x)
                for (size_t y = x + 1; y < nodes.size(); ++y) LG.add_edge(nodes[x], nodes[y]);
        TreeDecomp td = tree_decomposition(LG);
        tw = td.tw;
        for (int v : order_from_decomposition(td)) perm.push_back(nodeinfo[v][2]);
    } else if (perm.empty() && ncontr > 0) {
        return fail(QTN_EINVAL, "qtn_order_treewidth: empty line graph");
    }
    if ((int)perm.size() != ncontr) return fail(QTN_EINVAL, "qtn_order_treewidth: order is not a permutation");
    for (int k = 0; k < ncontr; ++k) perm_out[k] = perm[k] + 1;
    if (tw_out) *tw_out = tw;
    return QTN_OK;
}

int graph_treewidth(int nv, int ne, const int32_t* edges, int32_t* tw_out, int32_t* ordering_out) {
    if (nv <= 0) return fail(QTN_EINVAL, "qtn_graph_treewidth: empty graph");
    SGraph G(nv);
    for (int e = 0; e < ne; ++e) {
        int a = edges[2 * e] - 1, b = edges[2 * e + 1] - 1;
        if (a < 0 || b < 0 || a >= nv || b >= nv) return fail(QTN_EINVAL, "qtn_graph_treewidth: vertex out of range");
        G.add_edge(a, b);
    }
    TreeDecomp td = tree_decomposition(G);
    if (tw_out) *tw_out = td.tw;
    if (ordering_out) for (int i = 0; i < nv; ++i) ordering_out[i] = td.ordering[i] + 1;
    return QTN_OK;
}

// ---------------------------------------------------------------------------
// contract_order (src/contract.jl:184-235) with check_contraction! (:96-160)
// and getBuildCost (:68-86).  Costs are Int64; `Inf` is represented by INT64_MAX.
// ---------------------------------------------------------------------------
struct CObj {
    std::vector<char> leg, tf;
    std::vector<int> seq;
    int64_t cost;
};

int order_exhaustive(int nt, const int32_t* ranks, const int32_t* const* labels, int nlabels,
                     const int64_t* legdims, int32_t* seq_out, int32_t* nseq_out, int64_t* cost_out) {
    if (nt < 2) return fail(QTN_EINVAL, "qtn_order_exhaustive: need at least two tensors");
    const int64_t INF = INT64_MAX;
    std::vector<std::vector<CObj>> S(nt);
    std::vector<std::vector<char>> isnew(nt);
    for (int i = 0; i < nt; ++i) {
        CObj o;
        o.leg.assign(nlabels, 0);
        for (int j = 0; j < ranks[i]; ++j) {
            int l = std::abs(labels[i][j]);
            if (l < 1 || l > nlabels) return fail(QTN_EINVAL, "qtn_order_exhaustive: label out of range");
            o.leg[l - 1] = 1;
        }
        o.tf.assign(nt, 0);
        o.tf[i] = 1;
        o.cost = 0;
        S[0].push_back(std::move(o));
    }
    isnew[0].assign(nt, 1);
    int64_t mu_old = 0, mu_cap = 1, mu_next = INF;
    std::vector<char> common(nlabels), freel(nlabels), tin(nt);
    int guard = 0;
    while (S[nt - 1].empty()) {
        for (int c = 2; c <= nt; ++c) {
            for (int d = 1; d <= c / 2; ++d) {
                int a = d, b = c - d;
                auto& Sa = S[a - 1];
                auto& Sb = S[b - 1];
                auto& Sab = S[a + b - 1];
                for (size_t i = 0; i < Sa.size(); ++i) {
                    for (size_t j = 0; j < Sb.size(); ++j) {
                        const CObj& Ta = Sa[i];
                        const CObj& Tb = Sb[j];
                        bool overlap = false;
                        for (int t = 0; t < nt; ++t) if (Ta.tf[t] & Tb.tf[t]) { overlap = true; break; }
                        if (overlap) continue;
                        bool anyc = false;
                        for (int l = 0; l < nlabels; ++l) {
                            common[l] = Ta.leg[l] & Tb.leg[l];
                            freel[l] = Ta.leg[l] ^ Tb.leg[l];
                            anyc |= (bool)common[l];
                        }
                        if (!anyc) continue;
                        // getBuildCost
                        int64_t nc = 1;
                        for (int l = 0; l < nlabels; ++l) if (freel[l] | common[l]) nc *= legdims[l];
                        nc += Ta.cost + Tb.cost;
                        bool ok = true;
                        int64_t ret = nc;
                        if (nc > mu_cap) ok = false;
                        else if (!(isnew[a - 1][i] || isnew[b - 1][j]) && nc <= mu_old) { ret = INF; ok = false; }
                        if (!ok) { mu_next = std::min(mu_next, ret); continue; }
                        for (int t = 0; t < nt; ++t) tin[t] = Ta.tf[t] | Tb.tf[t];
                        int objptr = -1;
                        for (size_t p = 0; p < Sab.size(); ++p) if (Sab[p].tf == tin) { objptr = (int)p; break; }
                        if (objptr >= 0 && !(Sab[objptr].cost > nc)) continue;
                        CObj o;
                        o.leg = freel;
                        o.tf = tin;
                        o.seq = Ta.seq;
                        o.seq.insert(o.seq.end(), Tb.seq.begin(), Tb.seq.end());
                        for (int l = 0; l < nlabels; ++l) if (common[l]) o.seq.push_back(l + 1);
                        o.cost = nc;
                        if (objptr < 0) { Sab.push_back(std::move(o)); isnew[a + b - 1].push_back(1); }
                        else { Sab[objptr] = std::move(o); isnew[a + b - 1][objptr] = 1; }
                    }
                }
            }
        }
        mu_old = mu_cap;
        mu_cap = mu_next;
        mu_next = INF;
        for (auto& f : isnew) std::fill(f.begin(), f.end(), 0);
        if (mu_cap == INF && S[nt - 1].empty()) {
            if (++guard > 1) return fail(QTN_EINVAL, "qtn_order_exhaustive: network is disconnected (reference loops forever)");
        }
    }
    const CObj& r = S[nt - 1][0];
    *nseq_out = (int)r.seq.size();
    for (size_t i = 0; i < r.seq.size(); ++i) seq_out[i] = r.seq[i];
    if (cost_out) *cost_out = r.cost;
    return QTN_OK;
}

}  // namespace qtn
