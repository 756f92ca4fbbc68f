"""``tensor_circuit!`` and ``decompose!`` (src/tensor_circuit.jl:14-80, src/decompose.jl:6-52):
symbolic network growth on the host; the small gate-splitting SVDs go through
``qtn_svd_trunc`` like every other SVD of the path."""
import numpy as np

from .gates import CircuitGate
from .svd import operator_chain
from .tensor_network import Summation, Tensor


def decompose(cg):
    """``decompose!(cg)``: M-qubit gate -> chain of M tensors by sequential SVD."""
    M = len(cg.iwire)
    if not M > 1:
        raise ValueError("Only decompose Circuit Gates that apply to multiple wires")
    # src/decompose.jl:17-48: one device call for the reshape / permutedims / sequential-SVD chain
    sites = operator_chain(np.array(cg.matrix), M, entry="qtn_decompose")
    tensors = [Tensor(t) for t in sites]
    wires = [cg.iwire[i] for i in range(M)]
    bonds = [0, 3] + [4] * (M - 2)
    return tensors, bonds, wires


def tensor_circuit(psi, cgc, is_decompose=False):
    """``tensor_circuit!(psi, cgc; is_decompose)``.  The non-decomposed path contracts the
    state with gate legs 1..M (row bits) exactly as src/tensor_circuit.jl:44-51 does."""
    if isinstance(cgc, CircuitGate):
        cgc = [cgc]
    for cg in cgc:
        M = len(cg.iwire)
        if not cg.req_wires() <= len(psi.openidx):
            raise AssertionError("gate needs more wires than the network has open legs")
        if M > 1 and is_decompose:
            ts, cs, ws = decompose(cg)
            for i, (t, c, w) in enumerate(zip(ts, cs, ws), 1):
                psi.tensors.append(t)
                nt = len(psi.tensors)
                psi.contractions.append(Summation([psi.openidx[w - 1], (nt, 2 if i == 1 else 3)]))
                if c > 0:
                    psi.contractions.append(Summation([(nt - 1, c), (nt, 1)]))
                psi.openidx[w - 1] = (nt, 1 if i == 1 else 2)
            continue
        # column-major storage: the C ABI then takes the buffer as is (no per-call conversion copy)
        psi.tensors.append(Tensor(np.asfortranarray(np.reshape(cg.matrix, (2,) * (2 * M), order="F"))))
        nt = len(psi.tensors)
        for i, w in enumerate(cg.iwire, 1):
            psi.contractions.append(Summation([psi.openidx[w - 1], (nt, i)]))
            psi.openidx[w - 1] = (nt, M + i)
