// Host planner: ncon label walk -> pairwise steps with gather/scatter offset tables,
// slice handling, arena layout, cost model and the deterministic slice chooser.
//
// The walk restates TensorOperations.ncon 3.1.0 as called from
// src/contract.jl:257, 263 (labels processed in order; the two groups holding a
// label are contracted over ALL labels they share; output axes follow the
// negative labels -1, -2, ...).  Unlike the reference's TTGT, no operand is
// permuted: every step is one gather-GEMM whose operands are addressed through
// additive offset tables (row offset + k offset), and whose result is written in
// the layout the next step / the caller wants.
#include <algorithm>
#include <cmath>
#include <map>
#include <memory>
#include <set>

#include "qtn_internal.h"

namespace qtn {

static const int64_t kMaxLo = 4096;
static const int kNumSM = 148;

OffTable make_table(std::vector<int64_t>& tables, const std::vector<int64_t>& extents,
                    const std::vector<int64_t>& strides, int64_t max_lo) {
    OffTable t;
    size_t nm = extents.size();
    t.n = 1;
    for (auto e : extents) t.n *= e;
    size_t j = 0;
    t.L = 1;
    while (j < nm && t.L * extents[j] <= max_lo) { t.L *= extents[j]; ++j; }
    if (j == 0 && nm > 0) { t.L = extents[0]; j = 1; }
    t.lo = (int64_t)tables.size();
    tables.resize(tables.size() + t.L, 0);
    for (int64_t i = 0; i < t.L; ++i) {
        int64_t r = i, off = 0;
        for (size_t q = 0; q < j; ++q) { off += (r % extents[q]) * strides[q]; r /= extents[q]; }
        tables[t.lo + i] = off;
    }
    int64_t nh = t.n / t.L;
    t.hi = (int64_t)tables.size();
    tables.resize(tables.size() + nh, 0);
    for (int64_t i = 0; i < nh; ++i) {
        int64_t r = i, off = 0;
        for (size_t q = j; q < nm; ++q) { off += (r % extents[q]) * strides[q]; r /= extents[q]; }
        tables[t.hi + i] = off;
    }
    return t;
}

int materialize_tables(Plan* p) {
    if (p->tables_ready) return QTN_OK;
    double total = 0;
    for (auto& sp : p->table_specs) {
        double n = 1;
        for (auto e : sp.extents) n *= (double)e;
        total += 2.0 * std::sqrt(n) + 8192.0;
    }
    if (total > 1.5e9) return fail(QTN_ENOMEM, "plan needs ~%.3g offset-table entries; slice the contraction", total);
    p->tables.clear();
    p->tables.push_back(0);
    auto fill = [&](OffTable& t) {
        if (t.spec < 0) return;
        const TableSpec& sp = p->table_specs[t.spec];
        int64_t max_lo = std::max<int64_t>(kMaxLo, (int64_t)std::sqrt((double)t.n));
        OffTable m = make_table(p->tables, sp.extents, sp.strides, max_lo);
        t.L = m.L; t.lo = m.lo; t.hi = m.hi;
    };
    for (auto& s : p->steps) { fill(s.a_row); fill(s.a_k); fill(s.b_k); fill(s.b_col); fill(s.c_row); fill(s.c_col); }
    p->tables_ready = true;
    return QTN_OK;
}

namespace {

int64_t prod(const std::vector<int64_t>& v) {
    int64_t p = 1;
    for (auto x : v) p *= x;
    return p;
}

struct Parsed {
    int nt;
    std::vector<std::vector<int>> labels;       // per input, leg order
    std::vector<std::vector<int64_t>> dims;
    std::map<int, int64_t> ldim;                 // label -> extent
    std::vector<int> order;                      // positive labels to walk
};

int parse(int nt, const int32_t* ranks, const int64_t* const* dims, const int32_t* const* labels,
          const int32_t* order, int norder, Parsed& P) {
    if (nt < 1) return fail(QTN_EINVAL, "contraction needs at least one tensor");
    P.nt = nt;
    std::map<int, int> count;
    for (int i = 0; i < nt; ++i) {
        if (ranks[i] < 0 || ranks[i] > 60) return fail(QTN_EINVAL, "tensor %d: unsupported rank %d", i + 1, ranks[i]);
        P.labels.emplace_back(labels[i], labels[i] + ranks[i]);
        P.dims.emplace_back(dims[i], dims[i] + ranks[i]);
        for (int j = 0; j < ranks[i]; ++j) {
            int l = labels[i][j];
            if (l == 0) return fail(QTN_EINVAL, "tensor %d leg %d carries no label", i + 1, j + 1);
            if (dims[i][j] < 1) return fail(QTN_EINVAL, "tensor %d leg %d has extent < 1", i + 1, j + 1);
            count[l]++;
            auto it = P.ldim.find(l);
            if (it == P.ldim.end()) P.ldim[l] = dims[i][j];
            else if (it->second != dims[i][j])
                return fail(QTN_EINVAL, "label %d joins legs of different extent (%lld vs %lld)", l,
                            (long long)it->second, (long long)dims[i][j]);
        }
    }
    for (auto& kv : count) {
        if (kv.first > 0 && kv.second != 2) return fail(QTN_EINVAL, "contracted label %d appears %d times (must be 2)", kv.first, kv.second);
        if (kv.first < 0 && kv.second != 1) return fail(QTN_EINVAL, "open label %d appears %d times (must be 1)", kv.first, kv.second);
    }
    if (order) {
        std::set<int> seen;
        for (int i = 0; i < norder; ++i) {
            if (order[i] <= 0 || !count.count(order[i])) return fail(QTN_EINVAL, "order entry %d is not a contracted label", order[i]);
            seen.insert(order[i]);
            P.order.push_back(order[i]);
        }
        for (auto& kv : count) if (kv.first > 0 && !seen.count(kv.first)) P.order.push_back(kv.first);  // leftovers ascending
    } else {
        for (auto& kv : count) if (kv.first > 0) P.order.push_back(kv.first);  // std::map: ascending
    }
    return QTN_OK;
}

// Symbolic tree: nodes' label lists (traces removed) and merge steps.
struct SymTree {
    std::vector<std::vector<int>> nodes;
    std::vector<std::array<int, 3>> steps;  // a, b, out
    std::vector<std::vector<int>> shared;
};

SymTree sym_tree(const Parsed& P) {
    SymTree T;
    for (auto& lab : P.labels) {
        std::vector<int> l;
        for (int x : lab) if (std::count(lab.begin(), lab.end(), x) == 1) l.push_back(x);
        T.nodes.push_back(l);
    }
    std::vector<int> alive(P.nt);
    for (int i = 0; i < P.nt; ++i) alive[i] = i;
    auto merge = [&](int a, int b) {
        const auto la = T.nodes[a], lb = T.nodes[b];
        std::vector<int> sh, out;
        for (int l : la) if (std::find(lb.begin(), lb.end(), l) != lb.end()) sh.push_back(l);
        for (int l : la) if (std::find(sh.begin(), sh.end(), l) == sh.end()) out.push_back(l);
        for (int l : lb) if (std::find(sh.begin(), sh.end(), l) == sh.end()) out.push_back(l);
        T.nodes.push_back(out);
        int o = (int)T.nodes.size() - 1;
        T.steps.push_back({a, b, o});
        T.shared.push_back(sh);
        *std::find(alive.begin(), alive.end(), a) = o;
        alive.erase(std::find(alive.begin(), alive.end(), b));
    };
    // label -> current holders, maintained incrementally
    for (int lab : P.order) {
        int h[2], nh = 0;
        for (int n : alive) {
            const auto& ln = T.nodes[n];
            if (std::find(ln.begin(), ln.end(), lab) != ln.end()) { if (nh < 2) h[nh] = n; ++nh; }
        }
        if (nh == 2) merge(h[0], h[1]);
    }
    while (alive.size() > 1) merge(alive[0], alive[1]);
    return T;
}

struct Cost {
    double flops = 0, bytes = 0;
    unsigned __int128 flops_exact = 0;  // the slice chooser compares exact integers (oracle: Python ints)
    int64_t mx = 1;
};

Cost sym_cost(const SymTree& T, int ninputs, const std::map<int, int64_t>& ldim, const std::set<int>& sliced) {
    auto size = [&](const std::vector<int>& labs) {
        int64_t s = 1;
        for (int l : labs) if (!sliced.count(l)) s *= ldim.at(l);
        return s;
    };
    Cost c;
    for (int i = 0; i < ninputs; ++i) c.mx = std::max(c.mx, size(T.nodes[i]));
    for (size_t s = 0; s < T.steps.size(); ++s) {
        int64_t K = size(T.shared[s]);
        int64_t M = size(T.nodes[T.steps[s][0]]) / K;
        int64_t N = size(T.nodes[T.steps[s][1]]) / K;
        c.flops += 8.0 * (double)M * (double)N * (double)K;
        c.flops_exact += (unsigned __int128)8 * (unsigned __int128)M * (unsigned __int128)N * (unsigned __int128)K;
        c.bytes += 16.0 * ((double)M * K + (double)K * N + (double)M * N);
        c.mx = std::max(c.mx, M * N);
    }
    return c;
}

// Simple first-fit arena allocator over element offsets (256-byte granules).
struct Arena {
    std::vector<std::pair<int64_t, int64_t>> free_;  // (off, size), sorted by off
    int64_t end = 0;
    static int64_t round(int64_t n) { return (n + 15) / 16 * 16; }
    int64_t alloc(int64_t n) {
        n = round(n);
        for (size_t i = 0; i < free_.size(); ++i)
            if (free_[i].second >= n) {
                int64_t off = free_[i].first;
                if (free_[i].second == n) free_.erase(free_.begin() + i);
                else { free_[i].first += n; free_[i].second -= n; }
                return off;
            }
        if (!free_.empty() && free_.back().first + free_.back().second == end) {  // grow the tail block
            int64_t off = free_.back().first;
            end = off + n;
            free_.pop_back();
            return off;
        }
        int64_t off = end;
        end += n;
        return off;
    }
    void release(int64_t off, int64_t n) {
        n = round(n);
        auto it = std::lower_bound(free_.begin(), free_.end(), std::make_pair(off, (int64_t)0));
        it = free_.insert(it, {off, n});
        if (it + 1 != free_.end() && it->first + it->second == (it + 1)->first) { it->second += (it + 1)->second; free_.erase(it + 1); }
        if (it != free_.begin() && (it - 1)->first + (it - 1)->second == it->first) { (it - 1)->second += it->second; free_.erase(it); }
    }
};

}  // namespace

int choose_slices(int nt, const int32_t* ranks, const int64_t* const* dims, const int32_t* const* labels,
                  const int32_t* order, int norder, int max_log2, int64_t min_slices, int32_t* labels_out,
                  int32_t* nlabels_out) {
    Parsed P;
    int rc = parse(nt, ranks, dims, labels, order, norder, P);
    if (rc) return rc;
    if (max_log2 < 0 || max_log2 > 62) return fail(QTN_EINVAL, "qtn_choose_slices: max_log2_elems out of range");
    SymTree T = sym_tree(P);
    std::vector<int> sliced;
    std::set<int> sl;
    const int64_t limit = (int64_t)1 << max_log2;
    auto nslices = [&](const std::set<int>& s) { unsigned __int128 n = 1; for (int l : s) n *= (unsigned __int128)P.ldim[l]; return n; };
    auto size = [&](const std::vector<int>& labs) { int64_t s = 1; for (int l : labs) if (!sl.count(l)) s *= P.ldim[l]; return s; };
    while (true) {
        Cost c = sym_cost(T, P.nt, P.ldim, sl);
        if (c.mx <= limit && nslices(sl) >= (unsigned __int128)std::max<int64_t>(min_slices, 1)) break;
        std::set<int> cand;
        for (auto& labs : T.nodes)
            if (size(labs) == c.mx)
                for (int l : labs) if (l > 0 && !sl.count(l) && P.ldim[l] > 1) cand.insert(l);
        if (cand.empty()) break;
        bool have = false;
        unsigned __int128 bf = 0; int64_t bm = 0; int bl = 0;
        for (int l : cand) {  // ascending
            std::set<int> s2 = sl;
            s2.insert(l);
            Cost c2 = sym_cost(T, P.nt, P.ldim, s2);
            unsigned __int128 f = c2.flops_exact * nslices(s2);
            if (!have || f < bf || (f == bf && (c2.mx < bm || (c2.mx == bm && l < bl)))) { have = true; bf = f; bm = c2.mx; bl = l; }
        }
        sliced.push_back(bl);
        sl.insert(bl);
    }
    for (size_t i = 0; i < sliced.size(); ++i) labels_out[i] = sliced[i];
    *nlabels_out = (int)sliced.size();
    return QTN_OK;
}

int build_plan(int nt, const int32_t* ranks, const int64_t* const* dims, const int32_t* const* labels,
               const int32_t* order, int norder, const int32_t* slice_labels, int nslice, int dtype, Plan** out) {
    Parsed P;
    int rc = parse(nt, ranks, dims, labels, order, norder, P);
    if (rc) return rc;
    if (dtype != QTN_C128 && dtype != QTN_C64) return fail(QTN_EINVAL, "dtype must be QTN_C128 or QTN_C64");
    std::unique_ptr<Plan> plan(new Plan());
    Plan& pl = *plan;
    pl.dtype = dtype;
    pl.nt = nt;
    std::set<int> sl;
    for (int i = 0; i < nslice; ++i) {
        int l = slice_labels[i];
        if (l <= 0 || !P.ldim.count(l)) return fail(QTN_EINVAL, "slice label %d is not a contracted label", l);
        if (sl.count(l)) return fail(QTN_EINVAL, "slice label %d repeated", l);
        sl.insert(l);
        pl.slice_labels.push_back(l);
        pl.slice_dims.push_back(P.ldim[l]);
        pl.nslices *= P.ldim[l];
    }
    // ---- input nodes (sliced modes dropped, original strides kept) ------------
    std::vector<std::vector<int64_t>> node_strides;
    std::vector<std::vector<int>> full;  // per node: labels incl. sliced ones (drive the walk)
    int64_t in_off = 0;
    std::vector<int> cur(nt);  // current node standing for input i (after its trace step)
    for (int i = 0; i < nt; ++i) {
        Node n;
        n.is_input = true;
        n.input_index = i;
        n.offset = in_off;
        std::vector<int64_t> st;
        int64_t s = 1;
        for (size_t j = 0; j < P.labels[i].size(); ++j) {
            int l = P.labels[i][j];
            if (sl.count(l)) {
                int pos = (int)(std::find(pl.slice_labels.begin(), pl.slice_labels.end(), l) - pl.slice_labels.begin());
                n.slice_strides.emplace_back(pos, s);
                n.slice_dep = true;
            } else {
                n.labels.push_back(l);
                n.dims.push_back(P.dims[i][j]);
                st.push_back(s);
            }
            s *= P.dims[i][j];
        }
        n.numel = prod(n.dims);
        in_off += (s + 15) / 16 * 16;
        pl.nodes.push_back(n);
        node_strides.push_back(st);
        full.push_back(P.labels[i]);
        cur[i] = i;
    }
    pl.input_elems = in_off;
    auto spec_table = [&pl](const std::vector<int64_t>& e, const std::vector<int64_t>& st) {
        OffTable t;
        t.n = 1;
        for (auto x : e) t.n *= x;
        t.spec = (int)pl.table_specs.size();
        pl.table_specs.push_back({e, st});
        return t;
    };

    auto dense_strides = [](const std::vector<int64_t>& d) {
        std::vector<int64_t> s(d.size());
        int64_t p = 1;
        for (size_t i = 0; i < d.size(); ++i) { s[i] = p; p *= d[i]; }
        return s;
    };
    auto sel = [](const Node& n, const std::vector<int64_t>& st, const std::vector<int>& labs,
                  std::vector<int64_t>& ext, std::vector<int64_t>& str) {
        ext.clear(); str.clear();
        for (int l : labs) {
            size_t p = std::find(n.labels.begin(), n.labels.end(), l) - n.labels.begin();
            ext.push_back(n.dims[p]);
            str.push_back(st[p]);
        }
    };

    // ---- trace steps for labels repeated on one tensor --------------------------
    for (int i = 0; i < nt; ++i) {
        const Node& n = pl.nodes[i];
        std::vector<int> freel, tr;
        for (int l : n.labels) {
            int c = (int)std::count(n.labels.begin(), n.labels.end(), l);
            if (c == 1) freel.push_back(l);
            else if (std::find(tr.begin(), tr.end(), l) == tr.end()) tr.push_back(l);
        }
        if (tr.empty()) continue;
        Node o;
        o.labels = freel;
        std::vector<int64_t> fe, fs, te, ts;
        sel(n, node_strides[i], freel, fe, fs);
        o.dims = fe;
        o.numel = prod(fe);
        o.slice_dep = n.slice_dep;
        for (int l : tr) {
            int64_t sum = 0, e = 0;
            for (size_t p = 0; p < n.labels.size(); ++p) if (n.labels[p] == l) { sum += node_strides[i][p]; e = n.dims[p]; }
            te.push_back(e);
            ts.push_back(sum);
        }
        Step s;
        s.kind = STEP_TRACE;
        s.a = i;
        s.out = (int)pl.nodes.size();
        s.M = o.numel;
        s.K = prod(te);
        s.a_row = spec_table(fe, fs);
        s.a_k = spec_table(te, ts);
        pl.nodes.push_back(o);
        node_strides.push_back(dense_strides(o.dims));
        {
            std::vector<int> f;
            for (int l : full[i]) if (std::count(full[i].begin(), full[i].end(), l) == 1) f.push_back(l);
            for (int l : tr) if (sl.count(l)) return fail(QTN_EINVAL, "slice label %d is a self-contraction", l);
            full.push_back(f);
        }
        pl.steps.push_back(s);
        cur[i] = s.out;
    }

    // ---- pairwise walk ------------------------------------------------------------
    std::vector<int> alive(cur.begin(), cur.end());
    auto merge = [&](int a, int b) {
        std::vector<int> sh;
        for (int l : pl.nodes[a].labels)
            if (std::find(pl.nodes[b].labels.begin(), pl.nodes[b].labels.end(), l) != pl.nodes[b].labels.end()) sh.push_back(l);
        auto freeof = [&](int x) {
            std::vector<int> f;
            for (int l : pl.nodes[x].labels) if (std::find(sh.begin(), sh.end(), l) == sh.end()) f.push_back(l);
            return f;
        };
        int a0 = a;
        std::vector<int> fa = freeof(a), fb = freeof(b);
        auto vol = [&](const std::vector<int>& f) { int64_t v = 1; for (int l : f) v *= P.ldim[l]; return v; };
        if (vol(fb) > vol(fa)) { std::swap(a, b); std::swap(fa, fb); }  // larger free side becomes M
        // shared labels in a's layout order
        sh.clear();
        for (int l : pl.nodes[a].labels)
            if (std::find(pl.nodes[b].labels.begin(), pl.nodes[b].labels.end(), l) != pl.nodes[b].labels.end()) sh.push_back(l);
        Step s;
        s.kind = STEP_GEMM;
        s.a = a;
        s.b = b;
        s.out = (int)pl.nodes.size();
        std::vector<int64_t> e, st;
        sel(pl.nodes[a], node_strides[a], fa, e, st);
        s.M = prod(e);
        s.a_row = spec_table(e, st);
        std::vector<int64_t> me = e;
        sel(pl.nodes[a], node_strides[a], sh, e, st);
        s.K = prod(e);
        s.a_k = spec_table(e, st);
        sel(pl.nodes[b], node_strides[b], sh, e, st);
        s.b_k = spec_table(e, st);
        sel(pl.nodes[b], node_strides[b], fb, e, st);
        s.N = prod(e);
        s.b_col = spec_table(e, st);
        s.n_mlabels = (int)fa.size();
        {
            std::vector<int> f;
            for (int l : full[a]) if (std::find(full[b].begin(), full[b].end(), l) == full[b].end()) f.push_back(l);
            for (int l : full[b]) if (std::find(full[a].begin(), full[a].end(), l) == full[a].end()) f.push_back(l);
            full.push_back(f);
        }
        Node o;
        o.labels = fa;
        o.labels.insert(o.labels.end(), fb.begin(), fb.end());
        o.dims = me;
        o.dims.insert(o.dims.end(), e.begin(), e.end());
        o.numel = s.M * s.N;
        o.slice_dep = pl.nodes[a].slice_dep || pl.nodes[b].slice_dep;
        s.c_dense = true;
        pl.nodes.push_back(o);
        node_strides.push_back(dense_strides(o.dims));
        pl.steps.push_back(s);
        *std::find(alive.begin(), alive.end(), a0) = s.out;
        alive.erase(std::find(alive.begin(), alive.end(), a0 == a ? b : a));
    };
    // The walk is driven by the FULL label sets, so every slice runs the tree of the
    // un-sliced network; sliced labels are only absent from the operand layouts.
    for (int lab : P.order) {
        int h[2], nh = 0;
        for (int n : alive) {
            const auto& ln = full[n];
            if (std::find(ln.begin(), ln.end(), lab) != ln.end()) { if (nh < 2) h[nh] = n; ++nh; }
        }
        if (nh == 2) merge(h[0], h[1]);
    }
    while (alive.size() > 1) merge(alive[0], alive[1]);
    int last = alive[0];

    // ---- output layout: open labels sorted descending (-1 first = fastest) ---------
    std::vector<int> olabels = pl.nodes[last].labels;
    std::sort(olabels.begin(), olabels.end(), [](int x, int y) { return x > y; });
    for (int l : olabels) {
        if (l > 0) return fail(QTN_EINVAL, "internal: contracted label %d survived the walk", l);
        pl.out_dims.push_back(P.ldim[l]);
    }
    pl.out_numel = prod(pl.out_dims);
    std::vector<int64_t> ostr_by_label_pos(olabels.size());
    {
        std::vector<int64_t> ds = dense_strides(pl.out_dims);
        for (size_t i = 0; i < olabels.size(); ++i) ostr_by_label_pos[i] = ds[i];
    }
    auto out_stride = [&](int l) {
        size_t p = std::find(olabels.begin(), olabels.end(), l) - olabels.begin();
        return ostr_by_label_pos[p];
    };
    bool last_is_gemm = !pl.steps.empty() && pl.steps.back().out == last && pl.steps.back().kind == STEP_GEMM;
    if (last_is_gemm) {
        Step& s = pl.steps.back();
        Node& o = pl.nodes[last];
        // the final GEMM scatters straight into the caller's layout
        std::vector<int64_t> e, st;
        size_t nm = (size_t)s.n_mlabels;
        for (size_t q = 0; q < nm; ++q) { e.push_back(o.dims[q]); st.push_back(out_stride(o.labels[q])); }
        s.c_row = spec_table(e, st);
        e.clear(); st.clear();
        for (size_t q = nm; q < o.labels.size(); ++q) { e.push_back(o.dims[q]); st.push_back(out_stride(o.labels[q])); }
        s.c_col = spec_table(e, st);
        s.c_dense = false;
        s.final_step = true;
    } else {
        // single tensor (or trace-only): permute into the caller's layout
        const Node& n = pl.nodes[last];
        Step s;
        s.kind = STEP_PERMUTE;
        s.a = last;
        s.out = (int)pl.nodes.size();
        s.final_step = true;
        s.M = n.numel;
        std::vector<int64_t> e, sin, sout;
        // enumerate in OUTPUT order so writes are contiguous
        for (int l : olabels) {
            size_t p = std::find(n.labels.begin(), n.labels.end(), l) - n.labels.begin();
            e.push_back(n.dims[p]);
            sin.push_back(node_strides[last][p]);
            sout.push_back(out_stride(l));
        }
        s.a_row = spec_table(e, sin);
        s.c_row = spec_table(e, sout);
        Node o;
        o.labels = olabels;
        o.dims = pl.out_dims;
        o.numel = pl.out_numel;
        o.slice_dep = n.slice_dep;
        pl.nodes.push_back(o);
        node_strides.push_back(dense_strides(o.dims));
        pl.steps.push_back(s);
        last = s.out;
    }
    pl.final_node = last;

    // ---- invariance, arena, variants, cost --------------------------------------------
    std::vector<int> last_use(pl.nodes.size(), -1);
    for (size_t i = 0; i < pl.steps.size(); ++i) {
        Step& s = pl.steps[i];
        s.invariant = pl.nslices > 1 && !pl.nodes[s.out].slice_dep && !s.final_step;
        if (s.a >= 0) last_use[s.a] = (int)i;
        if (s.b >= 0) last_use[s.b] = (int)i;
    }
    int64_t pers = 0;
    for (auto& s : pl.steps)
        if (s.invariant) { Node& o = pl.nodes[s.out]; o.persistent = true; o.offset = pers; pers += Arena::round(o.numel); }
    // Small plans are latency-bound: give every intermediate its own buffer (no WAR hazards) so
    // the executor can run independent steps concurrently (DAG capture).  Large plans re-use memory.
    int64_t no_reuse = 0;
    for (auto& s : pl.steps) if (!s.invariant && !s.final_step) no_reuse += Arena::round(pl.nodes[s.out].numel);
    pl.dag = no_reuse <= ((int64_t)1 << 26);  // <= 1 GiB of intermediates
    Arena ar;
    for (size_t i = 0; i < pl.steps.size(); ++i) {
        Step& s = pl.steps[i];
        Node& o = pl.nodes[s.out];
        if (!s.invariant && !s.final_step) o.offset = pers + ar.alloc(o.numel);
        if (pl.dag) continue;
        for (int x : {s.a, s.b}) {
            if (x < 0) continue;
            Node& n = pl.nodes[x];
            if (!n.is_input && !n.persistent && last_use[x] == (int)i) ar.release(n.offset - pers, n.numel);
        }
    }
    pl.arena_elems = pers + ar.end;
    pl.launches_per_slice = 0;
    for (auto& s : pl.steps) {
        if (s.kind == STEP_GEMM) {
            int bm = 64, bn = (s.N <= 16) ? 8 : 64;
            s.variant = (s.N <= 16) ? 1 : 0;
            int64_t tiles = ((s.M + bm - 1) / bm) * ((s.N + bn - 1) / bn);
            s.split_k = 1;
            if (s.M * s.N <= 16 && s.K >= 2048) {
                s.variant = 2;  // closing dot products: streaming gather-dot, HBM-bound
            } else if (tiles < kNumSM && s.K >= 256) {
                int64_t want = (2 * kNumSM + tiles - 1) / tiles;
                int64_t maxs = s.K / 64;
                s.split_k = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(want, maxs), 1024));
            }
            pl.flops += 8.0 * (double)s.M * (double)s.N * (double)s.K;
            pl.bytes += 16.0 * ((double)s.M * s.K + (double)s.K * s.N + (double)s.M * s.N);
            pl.max_elems = std::max(pl.max_elems, s.M * s.N);
        }
        if (!s.invariant) pl.launches_per_slice += 1;
    }
    for (int i = 0; i < nt; ++i) pl.max_elems = std::max(pl.max_elems, pl.nodes[i].numel);
    *out = plan.release();
    return QTN_OK;
}

}  // namespace qtn
