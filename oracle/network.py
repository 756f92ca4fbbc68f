"""Oracle data model (test infrastructure): the reference's L1 containers.

Follows ``src/tensor.jl:7-22``, ``src/tensor_network.jl:7-35`` and
``src/helper.jl:6-27``.  A leg reference ``tensor => leg`` is the tuple
``(tensor, leg)``, both 1-based as in the reference.
"""
import numpy as np


class Tensor:
    """``struct Tensor; data::Array; end`` (src/tensor.jl:7-10)."""

    def __init__(self, data):
        self.data = np.asarray(data)

    @property
    def ndims(self):
        return self.data.ndim

    @property
    def size(self):
        return tuple(self.data.shape)

    def reshape(self, *dims):  # src/tensor.jl:14-15, column-major
        if len(dims) == 1 and isinstance(dims[0], (tuple, list)):
            dims = tuple(dims[0])
        return Tensor(np.reshape(self.data, dims, order="F"))

    def transpose(self):  # src/tensor.jl:18
        return Tensor(self.data.T)

    def isapprox(self, other):  # src/tensor.jl:20-22
        return bool(np.all(np.isclose(self.data, other.data, rtol=1.5e-8, atol=0)))


class Summation:
    """``struct Summation; idx::Vector{Pair}; end`` (src/tensor_network.jl:7-14)."""

    def __init__(self, idx):
        self.idx = [tuple(p) for p in idx]

    def __eq__(self, other):
        return isinstance(other, Summation) and self.idx == other.idx

    def __hash__(self):
        return hash(tuple(self.idx))

    def __repr__(self):
        return "Summation(%r)" % (self.idx,)


class Network:
    """``GeneralTensorNetwork`` (src/tensor_network.jl:26-35); copy is shallow."""

    def __init__(self, tensors, contractions, openidx):
        self.tensors = list(tensors)
        self.contractions = list(contractions)
        self.openidx = [tuple(p) for p in openidx]

    def copy(self):
        return type(self)(list(self.tensors), list(self.contractions), list(self.openidx))


def shift_summation(S, step):  # src/helper.jl:6-8
    return Summation([(S.idx[i][0] + step, S.idx[i][1]) for i in range(2)])


def shift_pair(P, step):  # src/helper.jl:15-17
    return (P[0] + step, P[1])


def is_power_two(i):  # src/helper.jl:24-27
    return i != 0 and (i & (i - 1)) == 0
