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
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <set>

#include "qtn_internal.h"

namespace qtn {

static const int64_t kMaxLo = 4096;
// QTN_SPLIT_SMALL=0 disables the split-K of small-output mid-size steps (A/B runs)
static bool split_small() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("QTN_SPLIT_SMALL"); v = e ? atoi(e) : 1; }
    return v != 0;
}
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
    // mixed-radix enumeration by doubling: the table of modes 0..q is e_q shifted copies of the table of modes 0..q-1
    // (no division / modulo per entry: this loop was 40 ms of a 60 ms one-shot contract on 2^20-row intermediates)
    auto enumerate = [&](int64_t base, size_t q0, size_t q1) {
        int64_t len = 1;
        if ((size_t)base >= tables.size()) return;   // empty section
        tables[base] = 0;
        for (size_t q = q0; q < q1; ++q) {
            for (int64_t d = 1; d < extents[q]; ++d) {
                const int64_t add = d * strides[q];
                int64_t* dst = tables.data() + base + d * len;
                const int64_t* src = tables.data() + base;
                for (int64_t i = 0; i < len; ++i) dst[i] = src[i] + add;
            }
            len *= extents[q];
        }
    };
    t.lo = (int64_t)tables.size();
    tables.resize(tables.size() + t.L, 0);
    enumerate(t.lo, 0, j);
    int64_t nh = t.n / t.L;
    t.hi = (int64_t)tables.size();
    tables.resize(tables.size() + nh, 0);
    enumerate(t.hi, j, nm);
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
    {   // one allocation for all tables (growing the vector table by table cost more than filling it)
        size_t entries = 1;
        auto size_of = [&](const OffTable& t) {
            if (t.spec < 0) return;
            const TableSpec& sp = p->table_specs[t.spec];
            const int64_t max_lo = std::max<int64_t>(kMaxLo, (int64_t)std::sqrt((double)t.n));
            int64_t L = 1, n = 1;
            size_t j = 0;
            for (auto e : sp.extents) n *= e;
            while (j < sp.extents.size() && L * sp.extents[j] <= max_lo) { L *= sp.extents[j]; ++j; }
            if (j == 0 && !sp.extents.empty()) L = sp.extents[0];
            entries += (size_t)L + (size_t)(L ? n / L : 0);
        };
        for (auto& s : p->steps) { size_of(s.a_row); size_of(s.a_k); size_of(s.b_k); size_of(s.b_col); size_of(s.c_row); size_of(s.c_col); }
        p->tables.reserve(entries);
    }
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

namespace {

// The greedy slice rule on a parsed network; returns total flops over all slices (exact) and the labels.
struct SliceChoice {
    std::vector<int> labels;
    unsigned __int128 nslices = 1;
    Cost per_slice;
};

SliceChoice choose_slices_sym(const Parsed& P, const SymTree& T, int max_log2, int64_t min_slices) {
    SliceChoice R;
    std::set<int> sl;
    const int64_t limit = (int64_t)1 << max_log2;
    auto nslices = [&](const std::set<int>& s) { unsigned __int128 n = 1; for (int l : s) n *= (unsigned __int128)P.ldim.at(l); return n; };
    auto size = [&](const std::vector<int>& labs) { int64_t s = 1; for (int l : labs) if (!sl.count(l)) s *= P.ldim.at(l); return s; };
    while (true) {
        Cost c = sym_cost(T, P.nt, P.ldim, sl);
        R.per_slice = c;
        R.nslices = nslices(sl);
        if (c.mx <= limit && R.nslices >= (unsigned __int128)std::max<int64_t>(min_slices, 1)) break;
        std::set<int> cand;
        for (auto& labs : T.nodes)
            if (size(labs) == c.mx)
                for (int l : labs) if (l > 0 && !sl.count(l) && P.ldim.at(l) > 1) cand.insert(l);
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
        R.labels.push_back(bl);
        sl.insert(bl);
    }
    return R;
}

}  // namespace

int choose_slices(int nt, const int32_t* ranks, const int64_t* const* dims, const int32_t* const* labels,
                  const int32_t* order, int norder, int max_log2, int64_t min_slices, int32_t* labels_out,
                  int32_t* nlabels_out) {
    Parsed P;
    int rc = parse(nt, ranks, dims, labels, order, norder, P);
    if (rc) return rc;
    if (max_log2 < 0 || max_log2 > 62) return fail(QTN_EINVAL, "qtn_choose_slices: max_log2_elems out of range");
    SymTree T = sym_tree(P);
    SliceChoice R = choose_slices_sym(P, T, max_log2, min_slices);
    for (size_t i = 0; i < R.labels.size(); ++i) labels_out[i] = R.labels[i];
    *nlabels_out = (int)R.labels.size();
    // the rule runs out of labels when the largest tensor carries only open or extent-1 labels: say so here instead
    // of letting the caller discover it as a failed allocation (the labels found so far are still returned)
    if (R.per_slice.mx > ((int64_t)1 << max_log2))
        return fail(QTN_EDOMAIN, "qtn_choose_slices: cannot slice below 2^%d elements: after %d labels the largest tensor still has %lld "
                    "(its remaining labels are open or of extent 1)", max_log2, (int)R.labels.size(), (long long)R.per_slice.mx);
    if (R.nslices < (unsigned __int128)std::max<int64_t>(min_slices, 1))
        return fail(QTN_EDOMAIN, "qtn_choose_slices: only %lld slices available, %lld requested", (long long)R.nslices, (long long)min_slices);
    return QTN_OK;
}

// ---- EXTENSION (no reference counterpart) -------------------------------------
// Randomised greedy search for a cheaper pairwise order than the reference's treewidth order.
// The reference order stays the default everywhere; this is what `optimize_contraction_order!`
// would be replaced by on request.  One trial: repeatedly merge the connected pair minimising
//   sign(c) log2(|c| + 1) - T * Gumbel,   c = size(out) - alpha * (size(a) + size(b)),
// (alpha, T drawn per trial; trial 0 is the noise-free alpha = 1 greedy).  Trials are ranked by
// log2(flops) + 0.5 * max(0, log2(max tensor) - max_log2); the best four trees are then refined by
// simulated annealing on tree rotations (AnnealTree below; slicing-aware when a memory target is
// given), and all finalists are re-costed exactly through the planner's own walk and the
// deterministic slice rule; the cheapest (flops + 5.7 flop/B * bytes, summed over slices) wins.
// Deterministic for a given (network, ntrials, seed); single-threaded.
namespace {

// FP64 tensor ceiling / HBM bandwidth of a B200 (37.1 TFLOP/s / 6.5 TB/s): flops one byte of traffic is worth.
constexpr double kBalanceFlopPerByte = 5.7;

struct SplitMix {
    uint64_t s;
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uni() { return ((next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }  // (0, 1)
};

struct Trial {
    std::vector<int> seq;   // one contracted label per merge, in merge order
    std::vector<std::array<int, 2>> merges;  // node ids (inputs 0..nt-1, the m-th merge creates nt + m)
    double log2_flops = 0;  // of the un-sliced tree
    double log2_mx = 0;
    double obj = 0;
};

struct SearchNet {
    std::vector<int> orig;                 // label index -> label
    std::vector<double> lw;                // log2 extent
    std::vector<std::vector<int>> nodes0;  // per input: sorted label indices (traced labels removed)
};

double lsize(const SearchNet& N, const std::vector<int>& labs) {
    double s = 0;
    for (int l : labs) s += N.lw[l];
    return s;
}

Trial greedy_trial(const SearchNet& N, double alpha, double temp, SplitMix& rng) {
    struct Cand { double key; int a, b; };
    auto cmp = [](const Cand& x, const Cand& y) { return x.key > y.key || (x.key == y.key && (x.a > y.a || (x.a == y.a && x.b > y.b))); };
    std::vector<std::vector<int>> nodes = N.nodes0;
    std::vector<double> ls(nodes.size());
    std::vector<char> alive(nodes.size(), 1);
    std::vector<std::array<int, 2>> holder(N.orig.size(), std::array<int, 2>{-1, -1});
    for (size_t i = 0; i < nodes.size(); ++i) {
        ls[i] = lsize(N, nodes[i]);
        for (int l : nodes[i]) (holder[l][0] < 0 ? holder[l][0] : holder[l][1]) = (int)i;
    }
    std::vector<Cand> heap;
    auto push = [&](int a, int b) {
        // symmetric difference size
        const auto &la = nodes[a], &lb = nodes[b];
        double lo = 0;
        size_t i = 0, j = 0;
        while (i < la.size() || j < lb.size()) {
            if (j == lb.size() || (i < la.size() && la[i] < lb[j])) lo += N.lw[la[i++]];
            else if (i == la.size() || lb[j] < la[i]) lo += N.lw[lb[j++]];
            else { ++i; ++j; }
        }
        double c = std::exp2(lo) - alpha * (std::exp2(ls[a]) + std::exp2(ls[b]));
        double key = c >= 0 ? std::log2(c + 1.0) : -std::log2(1.0 - c);
        if (temp > 0) key -= temp * -std::log(-std::log(rng.uni()));
        heap.push_back({key, std::min(a, b), std::max(a, b)});
        std::push_heap(heap.begin(), heap.end(), cmp);
    };
    {
        std::set<std::pair<int, int>> seen;
        for (size_t l = 0; l < holder.size(); ++l) {
            int a = holder[l][0], b = holder[l][1];
            if (N.orig[l] < 0 || a < 0 || b < 0 || a == b) continue;
            if (seen.insert({std::min(a, b), std::max(a, b)}).second) push(a, b);
        }
    }
    Trial R;
    double flops = 0, mx = 0;
    for (double x : ls) mx = std::max(mx, x);
    while (!heap.empty()) {
        std::pop_heap(heap.begin(), heap.end(), cmp);
        Cand c = heap.back();
        heap.pop_back();
        if (!alive[c.a] || !alive[c.b]) continue;
        const auto &la = nodes[c.a], &lb = nodes[c.b];
        std::vector<int> out;
        double lk = 0;
        int first_shared = -1;
        size_t i = 0, j = 0;
        while (i < la.size() || j < lb.size()) {
            if (j == lb.size() || (i < la.size() && la[i] < lb[j])) out.push_back(la[i++]);
            else if (i == la.size() || lb[j] < la[i]) out.push_back(lb[j++]);
            else { lk += N.lw[la[i]]; if (first_shared < 0) first_shared = la[i]; ++i; ++j; }
        }
        if (first_shared < 0) continue;  // cannot happen: candidates share a label
        R.seq.push_back(N.orig[first_shared]);
        R.merges.push_back({c.a, c.b});
        flops += std::exp2(ls[c.a] + ls[c.b] - lk);
        alive[c.a] = alive[c.b] = 0;
        int o = (int)nodes.size();
        nodes.push_back(out);
        ls.push_back(lsize(N, out));
        alive.push_back(1);
        mx = std::max(mx, ls[o]);
        std::vector<int> nb;
        for (int l : nodes[o]) {
            auto& h = holder[l];
            for (int q = 0; q < 2; ++q) if (h[q] == c.a || h[q] == c.b) h[q] = o;
            if (N.orig[l] < 0) continue;
            int other = h[0] == o ? h[1] : h[0];
            if (other >= 0 && other != o && std::find(nb.begin(), nb.end(), other) == nb.end()) nb.push_back(other);
        }
        for (int d : nb) push(o, d);
    }
    R.log2_flops = std::log2(std::max(flops, 1.0)) + 3.0;
    R.log2_mx = mx;
    return R;
}

// Binary contraction tree over label bitsets, refined by simulated annealing on tree rotations:
// at a node p = (a.b).c the alternatives (a.c).b and (b.c).a change one intermediate only, so a
// move is costed locally.  Step cost model: MNK + kByteWeight * (MK + KN + MN), i.e. flops plus
// operand traffic at the machine balance (FP64 tensor ceiling / HBM bandwidth ~ 5.7 flop per byte
// -> 16 B * 5.7 / 8 flop ~ 11 per complex element).  Sliced labels have weight zero; with a
// memory target, moves creating a tensor above it are rejected.
struct AnnealTree {
    const SearchNet* N = nullptr;
    int nleaf = 0, W = 0, root = -1;
    std::vector<int> parent, kid[2];
    std::vector<uint64_t> bits;     // node-major, W words
    std::vector<double> lw;         // label weight with sliced labels zeroed
    std::vector<double> sz;         // log2 size of each node under lw
    static constexpr double kByteWeight = kBalanceFlopPerByte * 16.0 / 8.0;  // per complex element, in units of 8 flop

    uint64_t* b(int n) { return &bits[(size_t)n * W]; }
    const uint64_t* b(int n) const { return &bits[(size_t)n * W]; }
    double wsum(const uint64_t* x) const {
        double s = 0;
        for (int w = 0; w < W; ++w) for (uint64_t v = x[w]; v; v &= v - 1) s += lw[w * 64 + __builtin_ctzll(v)];
        return s;
    }
    double wsum_and(const uint64_t* x, const uint64_t* y) const {
        double s = 0;
        for (int w = 0; w < W; ++w) for (uint64_t v = x[w] & y[w]; v; v &= v - 1) s += lw[w * 64 + __builtin_ctzll(v)];
        return s;
    }
    bool share(const uint64_t* x, const uint64_t* y) const {
        for (int w = 0; w < W; ++w) if (x[w] & y[w] & contr[w]) return true;
        return false;
    }
    std::vector<uint64_t> contr;  // contracted (positive) labels
    double step_cost(int x, int y, double so) const {
        double k = wsum_and(b(x), b(y));
        return std::exp2(sz[x] + sz[y] - k) + kByteWeight * (std::exp2(sz[x]) + std::exp2(sz[y]) + std::exp2(so));
    }
    double node_cost(int p) const { return step_cost(kid[0][p], kid[1][p], sz[p]); }
    double total_cost() const {
        double c = 0;
        for (int n = nleaf; n < (int)parent.size(); ++n) c += node_cost(n);
        return c;
    }
    double max_size() const {
        double m = 0;
        for (double x : sz) m = std::max(m, x);
        return m;
    }
    void resize_all() { for (int n = 0; n < (int)parent.size(); ++n) sz[n] = wsum(b(n)); }

    // Builds the tree from a greedy merge sequence (must connect everything).
    bool init(const SearchNet& net, const std::vector<std::array<int, 2>>& merges) {
        N = &net;
        nleaf = (int)net.nodes0.size();
        if ((int)merges.size() != nleaf - 1) return false;
        int L = (int)net.orig.size();
        W = (L + 63) / 64;
        int nn = 2 * nleaf - 1;
        parent.assign(nn, -1);
        kid[0].assign(nn, -1);
        kid[1].assign(nn, -1);
        bits.assign((size_t)nn * W, 0);
        contr.assign(W, 0);
        lw = net.lw;
        lw.resize((size_t)W * 64, 0.0);
        for (int l = 0; l < L; ++l) if (net.orig[l] > 0) contr[l / 64] |= 1ull << (l % 64);
        for (int i = 0; i < nleaf; ++i) for (int l : net.nodes0[i]) b(i)[l / 64] |= 1ull << (l % 64);
        for (int m = 0; m < nleaf - 1; ++m) {
            int o = nleaf + m, x = merges[m][0], y = merges[m][1];
            kid[0][o] = x; kid[1][o] = y;
            parent[x] = parent[y] = o;
            for (int w = 0; w < W; ++w) b(o)[w] = b(x)[w] ^ b(y)[w];
        }
        root = nn - 1;
        sz.assign(nn, 0.0);
        resize_all();
        return true;
    }

    // One Metropolis sweep of `moves` random rotations at temperature T; cap < 0: no size cap.
    void anneal(int64_t moves, double T, double cap, SplitMix& rng) {
        int nint = nleaf - 1;
        std::vector<uint64_t> nb(W);
        for (int64_t it = 0; it < moves; ++it) {
            uint64_t r = rng.next();
            int p = nleaf + (int)(r % (uint64_t)nint);
            int s = (int)((r >> 32) & 1), t = (int)((r >> 33) & 1);
            int l = kid[s][p], c = kid[1 - s][p];
            if (l < nleaf) continue;
            int x = kid[t][l], y = kid[1 - t][l];  // new: l = (x.c), p = (l.y)
            if (!share(b(x), b(c))) continue;
            for (int w = 0; w < W; ++w) nb[w] = b(x)[w] ^ b(c)[w];
            double ns = wsum(nb.data());
            if (cap >= 0 && ns > cap + 1e-9 && ns > sz[l]) continue;
            double old_cost = node_cost(l) + node_cost(p);
            double kxc = wsum_and(b(x), b(c));
            double kly = wsum_and(nb.data(), b(y));
            double new_cost = std::exp2(sz[x] + sz[c] - kxc) + kByteWeight * (std::exp2(sz[x]) + std::exp2(sz[c]) + std::exp2(ns)) +
                              std::exp2(ns + sz[y] - kly) + kByteWeight * (std::exp2(ns) + std::exp2(sz[y]) + std::exp2(sz[p]));
            double dE = std::log2(new_cost) - std::log2(old_cost);
            if (dE > 0 && (T <= 0 || rng.uni() >= std::exp2(-dE / T))) continue;
            kid[0][l] = x; kid[1][l] = c;
            parent[c] = l;
            kid[s][p] = l; kid[1 - s][p] = y;
            parent[y] = p;
            std::copy(nb.begin(), nb.end(), b(l));
            sz[l] = ns;
        }
    }

    // Adds sliced labels by the planner's rule (label on a largest tensor minimising the total cost) until
    // every tensor fits 2^cap; then drops sliced labels that are no longer needed.
    void fit_slices(double cap, std::vector<int>& sliced) {
        for (;;) {
            double mx = max_size();
            if (mx <= cap + 1e-9) break;
            std::vector<uint64_t> cand(W, 0);
            for (int n = 0; n < (int)parent.size(); ++n)
                if (sz[n] >= mx - 1e-9) for (int w = 0; w < W; ++w) cand[w] |= b(n)[w] & contr[w];
            int best = -1;
            double bc = 0, bm = 0;
            for (int w = 0; w < W; ++w)
                for (uint64_t v = cand[w]; v; v &= v - 1) {
                    int l = w * 64 + __builtin_ctzll(v);
                    if (lw[l] <= 0) continue;
                    double keep = lw[l];
                    lw[l] = 0;
                    resize_all();
                    double c = total_cost() * std::exp2(keep), m = max_size();
                    if (best < 0 || c < bc || (c == bc && m < bm)) { best = l; bc = c; bm = m; }
                    lw[l] = keep;
                }
            if (best < 0) { resize_all(); break; }
            lw[best] = 0;
            sliced.push_back(best);
            resize_all();
        }
    }
    void drop_unneeded_slices(double cap, std::vector<int>& sliced) {
        for (size_t i = sliced.size(); i-- > 0;) {
            int l = sliced[i];
            lw[l] = N->lw[l];
            resize_all();
            if (max_size() <= cap + 1e-9) sliced.erase(sliced.begin() + i);
            else lw[l] = 0;
        }
        resize_all();
    }

    // Label sequence of the tree (children before parents, costlier subtree first).
    std::vector<int> sequence() const {
        std::vector<int> seq, stack{root};
        std::vector<int> post;
        while (!stack.empty()) {
            int n = stack.back();
            stack.pop_back();
            if (n < nleaf) continue;
            post.push_back(n);
            stack.push_back(kid[0][n]);
            stack.push_back(kid[1][n]);
        }
        for (size_t i = post.size(); i-- > 0;) {
            int n = post[i];
            const uint64_t *x = b(kid[0][n]), *y = b(kid[1][n]);
            for (int w = 0; w < W; ++w) {
                uint64_t v = x[w] & y[w] & contr[w];
                if (v) { seq.push_back(N->orig[w * 64 + __builtin_ctzll(v)]); break; }
            }
        }
        return seq;
    }
};

}  // namespace

int order_search(int nt, const int32_t* ranks, const int64_t* const* dims, const int32_t* const* labels,
                 int ntrials, uint64_t seed, int max_log2, int32_t* order_out, int32_t* norder_out, double* cost_out) {
    Parsed P;
    int rc = parse(nt, ranks, dims, labels, nullptr, 0, P);
    if (rc) return rc;
    if (ntrials < 1) return fail(QTN_EINVAL, "qtn_order_search: ntrials must be >= 1");
    if (max_log2 > 62) return fail(QTN_EINVAL, "qtn_order_search: max_log2_elems out of range");
    SearchNet N;
    std::map<int, int> index;
    for (auto& kv : P.ldim) {
        index[kv.first] = (int)N.orig.size();
        N.orig.push_back(kv.first);
        N.lw.push_back(std::log2((double)kv.second));
    }
    for (auto& lab : P.labels) {
        std::vector<int> l;
        for (int x : lab) if (std::count(lab.begin(), lab.end(), x) == 1) l.push_back(index[x]);
        std::sort(l.begin(), l.end());
        N.nodes0.push_back(l);
    }
    SplitMix rng{seed ^ 0x51ED270B35A1F2C7ull};
    const size_t keep = 8;
    std::vector<Trial> best;
    for (int t = 0; t < ntrials; ++t) {
        double alpha = 1.0, temp = 0.0;
        if (t > 0) {
            alpha = 2.0 * rng.uni();
            temp = std::exp2(-7.0 + 8.0 * rng.uni());  // 2^-7 .. 2
        }
        Trial R = greedy_trial(N, alpha, temp, rng);
        R.obj = R.log2_flops + (max_log2 >= 0 ? 0.5 * std::max(0.0, R.log2_mx - (double)max_log2) : 0.0);
        size_t pos = 0;
        while (pos < best.size() && best[pos].obj <= R.obj) ++pos;
        if (pos < keep) {
            best.insert(best.begin() + pos, std::move(R));
            if (best.size() > keep) best.pop_back();
        }
    }
    // refinement of the best trees by annealing (connected networks only)
    std::vector<std::vector<int>> cands;
    for (auto& t : best) cands.push_back(t.seq);
    const int n_anneal = 4, n_steps = 64;          // trees refined, temperature steps per phase
    const double mv = 0.125, t_hi = 0.5, t_lo = 0.01;  // moves per step = mv * ntrials * (nt - 1); log2-cost temperatures
    const double cap = max_log2 >= 0 ? (double)max_log2 : -1.0;
    for (size_t i = 0; i < best.size() && (int)i < n_anneal; ++i) {
        AnnealTree A;
        if (nt < 3 || !A.init(N, best[i].merges)) continue;
        int64_t moves = std::max<int64_t>(64, (int64_t)(mv * ntrials * (nt - 1)));
        std::vector<int> sliced;
        for (int st = 0; st < n_steps; ++st) A.anneal(moves, t_hi * std::pow(t_lo / t_hi, (double)st / (n_steps - 1)), -1.0, rng);
        A.anneal(4 * moves, 0.0, -1.0, rng);
        if (cap >= 0) {
            for (int round = 0; round < 3; ++round) {
                A.fit_slices(cap, sliced);
                for (int st = 0; st < n_steps; ++st)
                    A.anneal(moves, 0.5 * t_hi * std::pow(t_lo / t_hi, (double)st / (n_steps - 1)), cap, rng);
                A.anneal(4 * moves, 0.0, cap, rng);
                A.drop_unneeded_slices(cap, sliced);
            }
        }
        cands.push_back(A.sequence());
    }
    // exact re-costing of the finalists through the planner's walk + slice rule
    bool have = false;
    double bt = 0;
    SliceChoice bc;
    std::vector<int> border;
    for (size_t i = 0; i < cands.size(); ++i) {
        Parsed Q = P;
        Q.order = cands[i];
        std::set<int> seen(Q.order.begin(), Q.order.end());
        for (auto& kv : P.ldim) if (kv.first > 0 && !seen.count(kv.first)) Q.order.push_back(kv.first);
        SymTree T = sym_tree(Q);
        SliceChoice c = max_log2 >= 0 ? choose_slices_sym(Q, T, max_log2, 1) : SliceChoice();
        if (max_log2 < 0) c.per_slice = sym_cost(T, Q.nt, Q.ldim, {});
        // ranked by flops + machine balance * bytes (the annealer's step model, summed exactly by the planner)
        double tot = ((double)c.per_slice.flops + kBalanceFlopPerByte * c.per_slice.bytes) * (double)c.nslices;
        if (!have || tot < bt) { have = true; bt = tot; bc = c; border = Q.order; }
    }
    for (size_t i = 0; i < border.size(); ++i) order_out[i] = border[i];
    *norder_out = (int)border.size();
    if (cost_out) {
        cost_out[0] = bc.per_slice.flops * (double)bc.nslices;
        cost_out[1] = bc.per_slice.flops;
        cost_out[2] = (double)bc.nslices;
        cost_out[3] = std::log2((double)bc.per_slice.mx);
    }
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
    // A sliced label that is a self-contraction (both legs on ONE tensor) needs no special case: both legs are
    // dropped from the node and each contributes its stride to the slice offset, so slice d reads the diagonal
    // T[.., d, .., d, ..] and the sum over the slices is the trace (tests/test_host_planner.py, test_gpu_contract.py).
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
        std::vector<int64_t> e, st;
        // Pre-permute of the smaller operand (the one TTGT transpose that is worth its traffic).  k is enumerated
        // in A's layout order, so A always has its stride-1 mode first in its row or k group; B need not: when
        // neither B's first k mode nor its first column mode has stride 1, every 16-byte element of a B tile
        // comes from a different sector, and B is re-read once per row tile.  If B is small against the step's
        // traffic it is first rewritten dense as (k in A's order, columns): one gather pass, coalesced writes.
        {
            std::vector<int64_t> ek, sk_, ec, sc_;
            sel(pl.nodes[b], node_strides[b], sh, ek, sk_);
            sel(pl.nodes[b], node_strides[b], fb, ec, sc_);
            auto first_stride = [](const std::vector<int64_t>& ex, const std::vector<int64_t>& sx) {
                for (size_t q = 0; q < ex.size(); ++q) if (ex[q] > 1) return sx[q];
                return (int64_t)1;
            };
            const int64_t Kb = prod(ek), Nb = prod(ec), Ma = vol(fa);
            const bool scattered = first_stride(ek, sk_) != 1 && first_stride(ec, sc_) != 1;
            const bool reread = Ma > 128;                                        // more than one row tile
            const bool cheap = Kb * Nb >= 65536 && 4 * Kb * Nb <= Ma * (Kb + Nb);  // permute traffic << step traffic
            static int prepermute = -1;
            if (prepermute < 0) { const char* ev = getenv("QTN_PREPERMUTE"); prepermute = ev ? atoi(ev) : 1; }
            // QTN_PREPERMUTE: 0 = never, 1 = rule above (default), 2 = whenever B is scattered (exercises the path in tests)
            if (prepermute && scattered && ((reread && cheap) || prepermute == 2)) {
                const Node& nb = pl.nodes[b];
                Step ps;
                ps.kind = STEP_PERMUTE;
                ps.a = b;
                ps.out = (int)pl.nodes.size();
                ps.M = nb.numel;
                Node o;
                o.labels = sh;
                o.labels.insert(o.labels.end(), fb.begin(), fb.end());
                std::vector<int64_t> sin;
                for (int l : o.labels) {
                    size_t q = std::find(nb.labels.begin(), nb.labels.end(), l) - nb.labels.begin();
                    o.dims.push_back(nb.dims[q]);
                    sin.push_back(node_strides[b][q]);
                }
                o.numel = nb.numel;
                o.slice_dep = nb.slice_dep;
                ps.a_row = spec_table(o.dims, sin);
                ps.c_row = spec_table(o.dims, dense_strides(o.dims));
                pl.nodes.push_back(o);
                node_strides.push_back(dense_strides(o.dims));
                full.push_back(full[b]);
                pl.steps.push_back(ps);
                if (a0 == b) a0 = ps.out;  // keep the alive-list bookkeeping below pointed at live nodes
                *std::find(alive.begin(), alive.end(), b) = ps.out;
                b = ps.out;
            }
        }
        Step s;
        s.kind = STEP_GEMM;
        s.a = a;
        s.b = b;
        s.out = (int)pl.nodes.size();
        sel(pl.nodes[a], node_strides[a], fa, e, st);
        s.M = prod(e);
        s.a_row = spec_table(e, st);
        auto min_stride = [](const std::vector<int64_t>& ex, const std::vector<int64_t>& sx) {
            int64_t m = INT64_MAX;
            for (size_t q = 0; q < ex.size(); ++q) if (ex[q] > 1) m = std::min(m, sx[q]);
            return m;
        };
        const int64_t a_row_min = min_stride(e, st);
        std::vector<int64_t> me = e;
        sel(pl.nodes[a], node_strides[a], sh, e, st);
        s.K = prod(e);
        s.a_k = spec_table(e, st);
        s.a_kmajor = min_stride(e, st) < a_row_min;
        sel(pl.nodes[b], node_strides[b], sh, e, st);
        s.b_k = spec_table(e, st);
        const int64_t b_k_min = min_stride(e, st);
        sel(pl.nodes[b], node_strides[b], fb, e, st);
        s.N = prod(e);
        s.b_col = spec_table(e, st);
        s.b_kmajor = b_k_min < min_stride(e, st);
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
            if (s.N > 16 && s.N <= 32 && s.M >= 1024) { s.variant = 3; bm = 128; bn = 32; }  // 128x32 tile: all columns in one CTA
            int64_t tiles = ((s.M + bm - 1) / bm) * ((s.N + bn - 1) / bn);
            s.split_k = 1;
            if (s.M * s.N <= 16 && s.K >= 2048) {
                s.variant = 2;  // closing dot products: streaming gather-dot, HBM-bound
            } else if (tiles < kNumSM && s.K >= 256) {
                int64_t want = (2 * kNumSM + tiles - 1) / tiles;
                int64_t maxs = s.K / 64;
                s.split_k = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(want, maxs), 1024));
            } else if (tiles < kNumSM / 2 && s.K >= 32 && s.M * s.N <= 65536 && split_small()) {
                // Mid-size steps of latency-bound plans (cfg 2: 256 x 128 x 128 on 8 tiles took 46 us = 16 dependent
                // k-iterations of ~3 us on 8 of 148 SMs): one 8-deep k-iteration per CTA; the output is small, so the
                // atomic epilogue (<= 64 K elements per split) is cheap.
                int64_t want = (2 * kNumSM + tiles - 1) / tiles;
                int64_t maxs = s.K / 8;
                s.split_k = (int)std::max<int64_t>(1, std::min<int64_t>(want, maxs));
            }
            pl.flops += 8.0 * (double)s.M * (double)s.N * (double)s.K;
            pl.bytes += 16.0 * ((double)s.M * s.K + (double)s.K * s.N + (double)s.M * s.N);
            pl.max_elems = std::max(pl.max_elems, s.M * s.N);
        }
        if (!s.invariant) pl.launches_per_slice += 1;
    }
    for (int i = 0; i < nt; ++i) pl.max_elems = std::max(pl.max_elems, pl.nodes[i].numel);
    if (getenv("QTN_PLAN_DEBUG")) {  // operand layouts of the big steps (diagnostics)
        auto show = [&](const char* nm, const OffTable& t) {
            const TableSpec& sp = pl.table_specs[t.spec];
            fprintf(stderr, "   %s:", nm);
            for (size_t q = 0; q < sp.extents.size(); ++q) fprintf(stderr, " %lldx%lld", (long long)sp.extents[q], (long long)sp.strides[q]);
            fprintf(stderr, "\n");
        };
        for (size_t i = 0; i < pl.steps.size(); ++i) {
            const Step& s = pl.steps[i];
            if (s.kind != STEP_GEMM || s.invariant || (double)s.M * s.N * s.K < 1e9) continue;
            fprintf(stderr, "step %zu M=%lld N=%lld K=%lld variant %d split %d akm %d bkm %d\n", i, (long long)s.M, (long long)s.N,
                    (long long)s.K, s.variant, s.split_k, (int)s.a_kmajor, (int)s.b_kmajor);
            show("a_row", s.a_row); show("a_k", s.a_k); show("b_k", s.b_k); show("b_col", s.b_col);
        }
    }
    *out = plan.release();
    return QTN_OK;
}

}  // namespace qtn
