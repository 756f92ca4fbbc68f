"""examples/amplitude.c: a plain-C host harness on the C ABI alone (network builder + order + contract).  Without a GPU
the host part must work and the compute call must fail loudly (no CPU fallback); on a B200 the amplitude must match
the oracle to 1e-10."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, to_oracle
from oracle import contract as oc


def _build(tmp_path):
    exe = str(tmp_path / "amplitude")
    libdir = os.path.join(ROOT, "qaintensor.jl_b200", "lib")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "amplitude.c"), "-L", libdir, "-lqaintensor_cuda",
                    "-Wl,-rpath," + libdir, "-o", exe], check=True)
    return exe


def _write_circuit(q, path, nq, depth, seed):
    net, gates, bits = q.circuits.cfg2_network(nq, depth, seed=seed)
    with open(path, "wb") as f:
        f.write(struct.pack("<ii", nq, len(gates)))
        for g in gates:
            f.write(struct.pack("<ii", *g.iwire))
            f.write(np.asfortranarray(np.asarray(g.matrix, dtype=np.complex128)).tobytes(order="F"))
        f.write(struct.pack("<%di" % nq, *[int(b) for b in bits]))
    return net


def test_c_harness_host_part_and_loud_failure_without_gpu(q, tmp_path):
    from qaintensor_b200 import _lib
    if _lib.device_count() > 0:
        pytest.skip("GPU present: covered by the gpu-marked test")
    exe = _build(tmp_path)
    net = _write_circuit(q, str(tmp_path / "c.bin"), 12, 8, 5)
    for method in ("treewidth", "search"):
        out = subprocess.run([exe, str(tmp_path / "c.bin"), method], capture_output=True, text=True)
        assert "network: %d tensors, %d contractions, 0 open legs" % (len(net.tensors), len(net.contractions)) in out.stdout
        assert "order ok (%s)" % method in out.stdout
        assert out.returncode == 3 and "qtn_net_contract" in out.stderr and "amplitude" not in out.stdout


@pytest.mark.gpu
def test_c_harness_amplitude_gpu(gpu, tmp_path):
    q = gpu
    exe = _build(tmp_path)
    net = _write_circuit(q, str(tmp_path / "c.bin"), 14, 10, 6)
    want = complex(oc.contract(to_oracle(net)))
    for args in (["treewidth"], ["search"], ["treewidth", "7"]):
        out = subprocess.run([exe, str(tmp_path / "c.bin")] + args, capture_output=True, text=True)
        assert out.returncode == 0, out.stdout + out.stderr
        re, im = (float(x) for x in out.stdout.strip().splitlines()[-1].split()[1:3])
        assert abs(complex(re, im) - want) < 1e-10 * abs(want)
