"""Reader of tests/golden/reference_inputs_r02.txt (format: make_reference_inputs.py) for the tests that compare
against tests/golden/reference_r02.json, the fixture julia/gen_golden.jl produces from the reference itself."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
INPUTS = os.path.join(HERE, "reference_inputs_r02.txt")
REFERENCE = os.path.join(HERE, "reference_r02.json")


def _cplx(tokens, dims):
    v = np.array(tokens, dtype=np.float64)
    return np.reshape(v[0::2] + 1j * v[1::2], dims, order="F")


def read_inputs(path=INPUTS):
    """-> (networks, matrices): networks[name] = (list of arrays or shapes, [((t1, l1), (t2, l2)), ...], with_data)."""
    nets, mats = {}, {}
    lines = open(path).read().splitlines()
    i = 0
    while i < len(lines):
        tok = lines[i].split()
        if not tok:
            i += 1
        elif tok[0] == "network":
            name, nt, nc, wd = tok[1], int(tok[2]), int(tok[3]), tok[4] == "1"
            tensors = []
            for t in range(nt):
                tt = lines[i + 1 + t].split()
                r = int(tt[1])
                dims = tuple(int(x) for x in tt[2:2 + r])
                tensors.append(_cplx(tt[2 + r:], dims) if wd else np.ones(dims, dtype=np.complex128))
            cons = []
            for c in range(nc):
                ct = lines[i + 1 + nt + c].split()
                cons.append(((int(ct[1]), int(ct[2])), (int(ct[3]), int(ct[4]))))
            nets[name] = (tensors, cons, wd)
            i += 1 + nt + nc
        elif tok[0] == "matrix":
            mats[tok[1]] = _cplx(tok[4:], (int(tok[2]), int(tok[3])))
            i += 1
        else:
            raise ValueError("unknown record %r" % tok[0])
    return nets, mats
