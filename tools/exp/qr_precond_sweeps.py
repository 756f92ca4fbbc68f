"""Sweep counts of cyclic one-sided Jacobi with and without QR preconditioning (numpy emulation).

Emulates the scalar round-robin scheme (n/2 disjoint pairs per round) of csrc/svd.cu on the CPU to see how many
sweeps a preconditioner R^H (A P = Q R) saves on the matrices the MPS path produces.
"""
import sys
import numpy as np
import scipy.linalg as sla


def jacobi_sweeps(A, tol=1e-15, max_sweeps=60):
    A = A.copy()
    m, n = A.shape
    assert n % 2 == 0
    idx = list(range(n))
    sweeps = 0
    nrm2 = np.linalg.norm(A) ** 2
    while sweeps < max_sweeps:
        rotated = 0
        order = idx[:]
        for _ in range(n - 1):
            p = np.array(order[: n // 2])
            q = np.array(order[n // 2:][::-1])
            ap, aq = A[:, p], A[:, q]
            alpha = np.einsum("ij,ij->j", ap.conj(), ap).real
            beta = np.einsum("ij,ij->j", aq.conj(), aq).real
            gam = np.einsum("ij,ij->j", ap.conj(), aq)
            ag = np.abs(gam)
            act = (ag > tol * np.sqrt(alpha * beta)) & (ag > 1e-30 * nrm2)
            rotated += int(act.sum())
            if act.any():
                ph = np.where(ag > 0, gam / np.where(ag > 0, ag, 1), 1)
                zeta = (beta - alpha) / (2 * np.where(act, ag, 1))
                t = np.sign(zeta) / (np.abs(zeta) + np.sqrt(1 + zeta * zeta))
                t = np.where(zeta == 0, 1.0, t)
                c = 1 / np.sqrt(1 + t * t)
                s = c * t
                c = np.where(act, c, 1.0)
                s = np.where(act, s, 0.0)
                np_ = c * ap - s * np.conj(ph) * aq
                nq = s * ph * ap + c * aq
                A[:, p], A[:, q] = np_, nq
            order = [order[0]] + [order[-1]] + order[1:-1]
        sweeps += 1
        if rotated == 0:
            break
    return sweeps, np.sort(np.linalg.norm(A, axis=0))[::-1]


def theta(chi, rng):
    A = rng.standard_normal((chi, 2, chi)) + 1j * rng.standard_normal((chi, 2, chi))
    B = rng.standard_normal((chi, 2, chi)) + 1j * rng.standard_normal((chi, 2, chi))
    # make them look like a canonical MPS with decaying Schmidt spectrum
    lam = np.exp(-np.linspace(0, 12, chi))
    A = A * lam[None, None, :]
    G = sla.qr(rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4)))[0].reshape(2, 2, 2, 2)
    th = np.einsum("abcd,lcm,mdr->labr", G, A, B)
    return th.reshape(2 * chi, 2 * chi)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    rng = np.random.default_rng(1)
    for name, M in (("gauss", rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))), ("theta", theta(n // 2, rng))):
        s_ref = np.linalg.svd(M, compute_uv=False)
        sw0, s0 = jacobi_sweeps(M)
        # sort columns by norm (de Rijk), then plain
        o = np.argsort(-np.linalg.norm(M, axis=0))
        sw1, s1 = jacobi_sweeps(M[:, o])
        Q, R = np.linalg.qr(M[:, o])
        sw2, s2 = jacobi_sweeps(R.conj().T)
        Qp, Rp, P = sla.qr(M, pivoting=True)
        sw3, s3 = jacobi_sweeps(Rp.conj().T)
        Q2, R2 = np.linalg.qr(R.conj().T)  # second QR (Drmac: L then QR again)
        sw4, s4 = jacobi_sweeps(R2.conj().T)
        err = lambda s: np.max(np.abs(s - s_ref)) / s_ref[0]
        print(f"{name} n={n}: plain {sw0} ({err(s0):.1e})  sorted {sw1} ({err(s1):.1e})  QR->R^H {sw2} ({err(s2):.1e})  "
              f"pivQR->R^H {sw3} ({err(s3):.1e})  QR,QR {sw4} ({err(s4):.1e})")


main()
