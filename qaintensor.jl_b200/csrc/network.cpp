// Host-side network builder behind the C ABI: the symbolic part of the reference's
// `GeneralTensorNetwork` (src/tensor_network.jl:26-33), `tensor_circuit!` (src/tensor_circuit.jl:14-52, the
// non-decomposed branch :44-51), `apply_MPO(psi, mpo, iwire)` (src/mpo.jl:232-252), `contract_rep`
// (src/contract.jl:39-60), `optimize_contraction_order!` (src/network2graph.jl:473-479) and `contract`
// (src/contract.jl:242-264, default-order branch), so that a C / C++ / Julia harness can go from gate matrices to
// amplitudes without any host code of its own.  All indices are 1-based as in the reference.  Tensor data is
// copied into the handle (ComplexF64, column-major); nothing here touches the device except qtn_net_contract,
// which goes through the same plan executor as qtn_contract.
#include <complex>
#include <cstring>
#include <memory>
#include <vector>

#include "qtn_internal.h"

using namespace qtn;

namespace {

typedef std::complex<double> cplx;

struct NetTensor {
    std::vector<int64_t> dims;
    std::vector<cplx> data;
};
struct Pair {
    int32_t t, l;
};

}  // namespace

struct qtn_net {
    std::vector<NetTensor> tensors;
    std::vector<std::array<Pair, 2>> contractions;
    std::vector<Pair> openidx;
};

namespace {

int64_t numel(const std::vector<int64_t>& d) {
    int64_t n = 1;
    for (auto x : d) n *= x;
    return n;
}

int check_leg(const qtn_net& n, Pair p, const char* what) {
    if (p.t < 1 || p.t > (int)n.tensors.size()) return fail(QTN_EINVAL, "%s: tensor index %d out of range", what, p.t);
    if (p.l < 1 || p.l > (int)n.tensors[p.t - 1].dims.size()) return fail(QTN_EINVAL, "%s: leg %d out of range for tensor %d", what, p.l, p.t);
    return QTN_OK;
}

// contract_rep (src/contract.jl:39-60): +k for contraction k, -(i + ncontractions) for open leg i.
int labels_of(const qtn_net& n, std::vector<std::vector<int32_t>>& lab) {
    lab.clear();
    for (auto& t : n.tensors) lab.emplace_back(t.dims.size(), 0);
    const int nc = (int)n.contractions.size();
    for (int k = 0; k < nc; ++k)
        for (const Pair& p : n.contractions[k]) {
            int rc = check_leg(n, p, "contraction");
            if (rc) return rc;
            lab[p.t - 1][p.l - 1] = k + 1;
        }
    for (int i = 0; i < (int)n.openidx.size(); ++i) {
        const Pair& p = n.openidx[i];
        int rc = check_leg(n, p, "open index");
        if (rc) return rc;
        if (lab[p.t - 1][p.l - 1] != 0) return fail(QTN_EINVAL, "open leg participates in a contraction");
        lab[p.t - 1][p.l - 1] = -(i + 1) - nc;
    }
    for (auto& l : lab)
        for (int32_t x : l) if (x == 0) return fail(QTN_EINVAL, "tensor leg without contraction or open index");
    return QTN_OK;
}

struct Marshal {
    std::vector<int32_t> ranks;
    std::vector<const int64_t*> dims;
    std::vector<const int32_t*> labels;
    std::vector<const void*> data;
    std::vector<std::vector<int32_t>> lab;
    int fill(const qtn_net& n) {
        int rc = labels_of(n, lab);
        if (rc) return rc;
        for (size_t i = 0; i < n.tensors.size(); ++i) {
            ranks.push_back((int32_t)n.tensors[i].dims.size());
            dims.push_back(n.tensors[i].dims.data());
            labels.push_back(lab[i].data());
            data.push_back(n.tensors[i].data.data());
        }
        return QTN_OK;
    }
};

}  // namespace

extern "C" {

int qtn_net_create(int32_t nt, const void* const* host_data, const int32_t* ranks, const int64_t* const* dims,
                   int32_t ncontr, const int32_t* pairs, int32_t nopen, const int32_t* openidx, qtn_net** net_out) {
    if (!net_out || nt < 0 || ncontr < 0 || nopen < 0 || (nt > 0 && (!host_data || !ranks || !dims)) ||
        (ncontr > 0 && !pairs) || (nopen > 0 && !openidx))
        return fail(QTN_EINVAL, "qtn_net_create: bad argument");
    std::unique_ptr<qtn_net> n(new qtn_net());
    for (int i = 0; i < nt; ++i) {
        if (ranks[i] < 0 || ranks[i] > 60) return fail(QTN_EINVAL, "tensor %d: unsupported rank %d", i + 1, ranks[i]);
        NetTensor t;
        t.dims.assign(dims[i], dims[i] + ranks[i]);
        for (auto d : t.dims) if (d < 1) return fail(QTN_EINVAL, "tensor %d has an extent < 1", i + 1);
        const cplx* src = static_cast<const cplx*>(host_data[i]);
        if (!src) return fail(QTN_EINVAL, "tensor %d: null data", i + 1);
        t.data.assign(src, src + numel(t.dims));
        n->tensors.push_back(std::move(t));
    }
    for (int k = 0; k < ncontr; ++k) {
        std::array<Pair, 2> c{{{pairs[4 * k], pairs[4 * k + 1]}, {pairs[4 * k + 2], pairs[4 * k + 3]}}};
        for (const Pair& p : c) { int rc = check_leg(*n, p, "contraction"); if (rc) return rc; }
        n->contractions.push_back(c);
    }
    for (int i = 0; i < nopen; ++i) {
        Pair p{openidx[2 * i], openidx[2 * i + 1]};
        int rc = check_leg(*n, p, "open index");
        if (rc) return rc;
        n->openidx.push_back(p);
    }
    *net_out = n.release();
    return QTN_OK;
}

int qtn_net_destroy(qtn_net* net) {
    delete net;
    return QTN_OK;
}

int qtn_net_sizes(const qtn_net* net, int32_t sizes[3]) {
    if (!net || !sizes) return fail(QTN_EINVAL, "qtn_net_sizes: null argument");
    sizes[0] = (int32_t)net->tensors.size();
    sizes[1] = (int32_t)net->contractions.size();
    sizes[2] = (int32_t)net->openidx.size();
    return QTN_OK;
}

int qtn_net_structure(const qtn_net* net, int32_t* pairs_out, int32_t* openidx_out) {
    if (!net) return fail(QTN_EINVAL, "qtn_net_structure: null network");
    if (pairs_out)
        for (size_t k = 0; k < net->contractions.size(); ++k) {
            pairs_out[4 * k] = net->contractions[k][0].t; pairs_out[4 * k + 1] = net->contractions[k][0].l;
            pairs_out[4 * k + 2] = net->contractions[k][1].t; pairs_out[4 * k + 3] = net->contractions[k][1].l;
        }
    if (openidx_out)
        for (size_t i = 0; i < net->openidx.size(); ++i) { openidx_out[2 * i] = net->openidx[i].t; openidx_out[2 * i + 1] = net->openidx[i].l; }
    return QTN_OK;
}

int qtn_net_tensor(const qtn_net* net, int32_t i, int32_t* rank_out, int64_t* dims_out, const void** data_out) {
    if (!net) return fail(QTN_EINVAL, "qtn_net_tensor: null network");
    if (i < 1 || i > (int)net->tensors.size()) return fail(QTN_EINVAL, "qtn_net_tensor: tensor index %d out of range", i);
    const NetTensor& t = net->tensors[i - 1];
    if (rank_out) *rank_out = (int32_t)t.dims.size();
    if (dims_out) for (size_t j = 0; j < t.dims.size(); ++j) dims_out[j] = t.dims[j];
    if (data_out) *data_out = t.data.data();
    return QTN_OK;
}

// tensor_circuit!(psi, cgc), is_decompose = false (src/tensor_circuit.jl:44-51): the gate matrix (2^M x 2^M,
// column-major) is reshaped to 2M legs of extent 2; open leg `wire` is joined to gate leg i, and gate leg
// M + i becomes the new open leg of that wire.
int qtn_net_tensor_circuit(qtn_net* net, int32_t ngates, const int32_t* nwires, const int32_t* wires,
                           const void* const* matrices) {
    if (!net || ngates < 0 || (ngates > 0 && (!nwires || !wires || !matrices))) return fail(QTN_EINVAL, "qtn_net_tensor_circuit: bad argument");
    // validate every gate before the first one is appended: a bad gate must not leave a half-extended network
    size_t w0 = 0;
    for (int g = 0; g < ngates; ++g) {
        const int M = nwires[g];
        if (M < 1 || M > 20) return fail(QTN_EINVAL, "gate %d: unsupported number of wires %d", g + 1, M);
        int need = 0;
        for (int i = 0; i < M; ++i) {
            int w = wires[w0 + i];
            if (w < 1) return fail(QTN_EINVAL, "Wires must be positive integers.");
            for (int j = 0; j < i; ++j) if (wires[w0 + j] == w) return fail(QTN_EINVAL, "Repeated wires are not valid.");
            need = std::max(need, w);
        }
        if (need > (int)net->openidx.size()) return fail(QTN_EINVAL, "gate needs more wires than the network has open legs");
        if (!matrices[g]) return fail(QTN_EINVAL, "gate %d: null matrix", g + 1);
        w0 += (size_t)M;
    }
    w0 = 0;
    for (int g = 0; g < ngates; ++g) {
        const int M = nwires[g];
        const cplx* src = static_cast<const cplx*>(matrices[g]);
        NetTensor t;
        t.dims.assign(2 * M, 2);
        t.data.assign(src, src + ((size_t)1 << (2 * M)));
        net->tensors.push_back(std::move(t));
        const int nt = (int)net->tensors.size();
        for (int i = 1; i <= M; ++i) {
            int w = wires[w0 + i - 1];
            net->contractions.push_back({{net->openidx[w - 1], Pair{nt, i}}});
            net->openidx[w - 1] = Pair{nt, M + i};
        }
        w0 += (size_t)M;
    }
    return QTN_OK;
}

// apply_MPO(psi, mpo::MPO, iwire) (src/mpo.jl:232-252): `op` is an operator network with 2M open legs in the
// MPO convention (openidx[1..M] = outputs, openidx[M+1..2M] = inputs, both listed from the LAST wire to the first,
// src/mpo.jl:88); returns a new network, psi and op are left untouched (as in the reference).
int qtn_net_apply_mpo(const qtn_net* psi, const qtn_net* op, int32_t nw, const int32_t* iwire, qtn_net** net_out) {
    if (!psi || !op || !iwire || !net_out || nw < 1) return fail(QTN_EINVAL, "qtn_net_apply_mpo: bad argument");
    if ((int)op->openidx.size() != 2 * nw) return fail(QTN_EINVAL, "operator network must have 2 * %d open legs, has %d", nw, (int)op->openidx.size());
    const int n = (int)psi->openidx.size();
    for (int i = 0; i < nw; ++i) {
        for (int j = 0; j < i; ++j) if (iwire[i] == iwire[j]) return fail(QTN_EINVAL, "Repeated wires are not valid.");
        if (iwire[i] < 1 || iwire[i] > n) return fail(QTN_EINVAL, "Wires must be integers between 1 and n (total number of qudits).");
    }
    const int step = (int)psi->tensors.size();
    std::unique_ptr<qtn_net> out(new qtn_net(*psi));
    for (auto& t : op->tensors) out->tensors.push_back(t);
    auto shift = [step](Pair p) { return Pair{p.t + step, p.l}; };
    for (auto& c : op->contractions) out->contractions.push_back({{shift(c[0]), shift(c[1])}});
    for (int i = 1; i <= nw; ++i) {
        int w = iwire[nw - i];
        out->contractions.push_back({{psi->openidx[w - 1], shift(op->openidx[i + nw - 1])}});
    }
    for (int i = 1; i <= nw; ++i) {
        int w = iwire[nw - i];
        out->openidx[w - 1] = shift(op->openidx[i - 1]);
    }
    *net_out = out.release();
    return QTN_OK;
}

// extend_MPO(mpo::MPO, iwire) (src/mpo.jl:122-157): an operator on M qubits (wires sorted descending) becomes one on
// the N = iwire[1] - iwire[M] + 1 qubits it spans by inserting identity "pipe" tensors delta(a, e) delta(b, c) of shape
// (bond, 2, 2, bond) on the wires in between.  Mutates `mpo` like the reference; the reference's error strings are kept.
int qtn_net_extend_mpo(qtn_net* mpo, int32_t nw, const int32_t* iwire) {
    if (!mpo || !iwire || nw < 1) return fail(QTN_EINVAL, "qtn_net_extend_mpo: bad argument");
    for (int i = 0; i < nw; ++i)
        for (int j = 0; j < i; ++j) if (iwire[i] == iwire[j]) return fail(QTN_EINVAL, "Repeated wires are not valid.");
    for (int i = 1; i < nw; ++i) if (iwire[i] > iwire[i - 1]) return fail(QTN_EINVAL, "Wires not sorted");
    for (int i = 0; i < nw; ++i) if (iwire[i] < 1) return fail(QTN_EINVAL, "Wires must be positive integers.");
    const int lo = iwire[nw - 1], hi = iwire[0];
    const int N = hi - lo + 1, M = nw;
    if ((int)mpo->tensors.size() != M) return fail(QTN_EINVAL, "MPO length does not match the wires");
    if (M == N) return fail(QTN_EINVAL, "MPO is already decomposed in N tensors");
    // qwire reversed = hi, hi-1, ..., lo; pipes = wires in [lo, hi] the operator does not act on, ascending
    for (int w = lo; w <= hi; ++w) {
        bool acts = false;
        for (int i = 0; i < nw; ++i) acts = acts || iwire[i] == w;
        if (acts) continue;
        const int ind = hi - w + 1;  // 1-based position of w in the reversed wire list
        if (ind < 2 || ind - 1 > (int)mpo->tensors.size()) return fail(QTN_EINVAL, "internal: pipe position out of range");
        const NetTensor& prev = mpo->tensors[ind - 2];
        if (prev.dims.empty()) return fail(QTN_EINVAL, "MPO tensor %d has no bond leg", ind - 1);
        const int64_t bond = prev.dims.back();
        NetTensor t;
        t.dims = {bond, 2, 2, bond};
        t.data.assign((size_t)(bond * 4 * bond), cplx(0.0, 0.0));
        for (int64_t a = 0; a < bond; ++a)
            for (int64_t b = 0; b < 2; ++b) t.data[(size_t)(a + bond * (b + 2 * (b + 2 * a)))] = cplx(1.0, 0.0);
        mpo->tensors.insert(mpo->tensors.begin() + (ind - 1), std::move(t));
    }
    for (int i = M; i < N; ++i) {
        mpo->contractions.push_back({{Pair{i, 4}, Pair{i + 1, 1}}});
        mpo->openidx.insert(mpo->openidx.begin(), Pair{i + 1, 2});
        if (i + 1 > (int)mpo->openidx.size()) return fail(QTN_EINVAL, "operator network has too few open legs for extend_MPO");
        mpo->openidx.insert(mpo->openidx.begin() + (i + 1), Pair{i + 1, 3});
    }
    return QTN_OK;
}

// EXTENSION (amplitude networks of BASELINE configs 2 and 3): closes every open leg w with the basis bra <bits[w]|.
int qtn_net_close(qtn_net* net, const int32_t* bits) {
    if (!net || (!bits && !net->openidx.empty())) return fail(QTN_EINVAL, "qtn_net_close: null argument");
    for (size_t w = 0; w < net->openidx.size(); ++w) {
        const Pair p = net->openidx[w];
        const int64_t d = net->tensors[p.t - 1].dims[p.l - 1];
        if (bits[w] < 0 || bits[w] >= d) return fail(QTN_EINVAL, "bit %d of wire %d out of range", bits[w], (int)w + 1);
        NetTensor t;
        t.dims = {d};
        t.data.assign((size_t)d, cplx(0.0, 0.0));
        t.data[bits[w]] = cplx(1.0, 0.0);
        net->tensors.push_back(std::move(t));
        net->contractions.push_back({{p, Pair{(int32_t)net->tensors.size(), 1}}});
    }
    net->openidx.clear();
    return QTN_OK;
}

// optimize_contraction_order!(net) (src/network2graph.jl:473-479): net.contractions = net.contractions[perm].
// method 0 = the reference's treewidth heuristic (bit-exact; the reference's two warnings about open indices are
// the caller's to print), 1 = EXTENSION qtn_order_search.
int qtn_net_optimize_order(qtn_net* net, int32_t method, int32_t ntrials, uint64_t seed, int32_t max_log2_elems) {
    if (!net) return fail(QTN_EINVAL, "qtn_net_optimize_order: null network");
    const int nc = (int)net->contractions.size();
    std::vector<int32_t> perm(std::max(nc, 1));
    if (method == 0) {
        std::vector<int32_t> pairs(4 * std::max(nc, 1));
        qtn_net_structure(net, pairs.data(), nullptr);
        int rc = qtn_order_treewidth((int32_t)net->tensors.size(), nc, pairs.data(), perm.data(), nullptr);
        if (rc) return rc;
    } else if (method == 1) {
        Marshal m;
        int rc = m.fill(*net);
        if (rc) return rc;
        int32_t n = 0;
        std::vector<int32_t> order;
        size_t cap = 1;
        for (auto r : m.ranks) cap += (size_t)r;
        order.resize(cap);
        rc = qtn_order_search((int32_t)net->tensors.size(), m.ranks.data(), m.dims.data(), m.labels.data(), ntrials, seed,
                              max_log2_elems, order.data(), &n, nullptr);
        if (rc) return rc;
        if (n != nc) return fail(QTN_EINVAL, "internal: searched order is not a permutation");
        for (int i = 0; i < nc; ++i) perm[i] = order[i];
    } else {
        return fail(QTN_EINVAL, "qtn_net_optimize_order: method must be 0 (treewidth) or 1 (search)");
    }
    std::vector<std::array<Pair, 2>> re;
    for (int i = 0; i < nc; ++i) re.push_back(net->contractions[perm[i] - 1]);
    net->contractions.swap(re);
    return QTN_OK;
}

// contract(net) (src/contract.jl:242-264, optimize = false): labels by contract_rep, default ascending order.
// max_log2_elems >= 0 (EXTENSION): sliced execution, as the keyword of the Python mirror's contract().
// dtype = QTN_C64: the tensors are rounded to ComplexF32 and host_out receives float pairs.
int qtn_net_contract(const qtn_net* net, int32_t dtype, int32_t max_log2_elems, void* host_out, int32_t* out_rank,
                     int64_t* out_dims) {
    QTN_API_GUARD();
    if (!net || !host_out) return fail(QTN_EINVAL, "qtn_net_contract: null argument");
    if (net->tensors.empty()) return fail(QTN_EINVAL, "contraction needs at least one tensor");
    Marshal m;
    int rc = m.fill(*net);
    if (rc) return rc;
    std::vector<std::vector<std::complex<float>>> f32;
    if (dtype == QTN_C64) {
        for (size_t i = 0; i < net->tensors.size(); ++i) {
            f32.emplace_back(net->tensors[i].data.begin(), net->tensors[i].data.end());
            m.data[i] = f32.back().data();
        }
    }
    const int nt = (int)net->tensors.size();
    if (max_log2_elems < 0)
        return qtn_contract(nt, m.data.data(), m.ranks.data(), m.dims.data(), m.labels.data(), nullptr, 0, dtype, host_out,
                            out_rank, out_dims);
    std::vector<int32_t> sl(std::max<size_t>(net->contractions.size(), 1));
    int32_t nsl = 0;
    rc = qtn_choose_slices(nt, m.ranks.data(), m.dims.data(), m.labels.data(), nullptr, 0, max_log2_elems, 1, sl.data(), &nsl);
    if (rc) return rc;
    qtn_plan* plan = nullptr;
    rc = qtn_plan_create(nt, m.ranks.data(), m.dims.data(), m.labels.data(), nullptr, 0, sl.data(), nsl, dtype, &plan);
    if (rc) return rc;
    int64_t info[8];
    qtn_plan_info(plan, info, nullptr);
    if (out_rank) *out_rank = (int32_t)info[2];
    if (out_dims) qtn_plan_out_dims(plan, out_dims);
    memset(host_out, 0, (size_t)info[3] * (dtype == QTN_C64 ? 8 : 16));
    rc = qtn_plan_execute_host(plan, m.data.data(), 0, info[1], host_out);
    qtn_plan_destroy(plan);
    return rc;
}

}  // extern "C"
