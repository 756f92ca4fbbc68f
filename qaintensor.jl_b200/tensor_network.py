"""Host containers mirroring the reference's data model.

``Tensor`` (src/tensor.jl:7-22), ``Summation`` / ``GeneralTensorNetwork``
(src/tensor_network.jl:7-35) and the helpers of src/helper.jl:6-27.  Leg references
``tensor => leg`` are ``(tensor, leg)`` tuples, 1-based like the reference.  Arrays keep
the reference's axis order (axis i of the Julia array is numpy axis i-1); the C ABI
receives them column-major.
"""
import numpy as np


class Tensor:
    def __init__(self, data):
        self.data = np.asarray(data)

    def ndims(self):
        return self.data.ndim

    def size(self):
        return tuple(self.data.shape)

    def reshape(self, *dims):
        if len(dims) == 1 and isinstance(dims[0], (tuple, list)):
            dims = tuple(dims[0])
        return Tensor(np.reshape(self.data, dims, order="F"))

    def transpose(self):
        return Tensor(self.data.T)

    def isapprox(self, other):
        return bool(np.all(np.isclose(self.data, other.data, rtol=np.sqrt(np.finfo(float).eps), atol=0.0)))


class Summation:
    def __init__(self, idx):
        self.idx = [(int(t), int(l)) for (t, l) in idx]

    def __eq__(self, other):
        return isinstance(other, Summation) and self.idx == other.idx

    def __hash__(self):
        return hash(tuple(self.idx))

    def __repr__(self):
        return "Summation(%s)" % ", ".join("%d=>%d" % p for p in self.idx)


class TensorNetwork:
    """Abstract base (src/tensor_network.jl:19)."""

    tensors: list
    contractions: list
    openidx: list


class GeneralTensorNetwork(TensorNetwork):
    def __init__(self, tensors, contractions, openidx):
        self.tensors = list(tensors)
        self.contractions = list(contractions)
        self.openidx = [(int(t), int(l)) for (t, l) in openidx]

    def copy(self):  # shallow, like Base.copy (src/tensor_network.jl:35)
        return GeneralTensorNetwork(list(self.tensors), list(self.contractions), list(self.openidx))


def shift_summation(S, step):
    return Summation([(t + step, l) for (t, l) in S.idx[:2]])


def shift_pair(P, step):
    return (P[0] + step, P[1])


def is_power_two(i):
    return i != 0 and (i & (i - 1)) == 0
