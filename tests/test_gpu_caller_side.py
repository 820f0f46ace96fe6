"""-m gpu: the steps either side of the solver in the caller's workflow -- structured quaternion products on
the device: congruence X^H F X (orthogonalisation before the eigensolver), C = X C' (after it), pairing fill.  The
reference's in-tree analogue is the residual check of its own test (test.cc:104-105: V^H M V with zgemm3m)."""
import ctypes

import numpy as np
import pytest

from oracle import zquatev_oracle as O

pytestmark = pytest.mark.gpu


def _left(M):
    """left half of a structured matrix, column-major, as a device tensor [cols][2 rows]"""
    import torch
    c = M.shape[1] // 2
    return torch.from_numpy(np.ascontiguousarray(M[:, :c].T)).cuda()


def _structured(n, m, seed):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((n, m)) + 1j * rng.standard_normal((n, m))
    b = rng.standard_normal((n, m)) + 1j * rng.standard_normal((n, m))
    return np.block([[a, -b.conj()], [b, a.conj()]])


@pytest.mark.parametrize("n,m", [(40, 40), (130, 77), (300, 300)])
def test_congruence_and_back_multiplication(n, m):
    import torch
    from zquatev_b200 import api
    L = api.lib()
    F = O.gen_sym(n, 5)                                   # structured Hermitian "Fock" matrix
    X = _structured(n, m, 6)                              # structured transformation (2n x 2m)
    Fd, Xd = _left(F), _left(X)
    out = torch.zeros((2 * m, 2 * m), dtype=torch.complex128, device="cuda")      # [cols][rows], ld 2m
    work = torch.zeros((m, 2 * n), dtype=torch.complex128, device="cuda")
    rc = L.zquatev_b200_congruence(n, m, Xd.data_ptr(), 2 * n, Fd.data_ptr(), 2 * n, out.data_ptr(), 2 * m, work.data_ptr(), None)
    assert rc == 0
    assert L.zquatev_b200_fill_pairing(m, m, out.data_ptr(), 2 * m, None) == 0
    torch.cuda.synchronize()
    got = out.cpu().numpy().T
    ref = X.conj().T @ F @ X
    tol = 1e-13 * np.linalg.norm(F, 2) * np.linalg.norm(X, 2) ** 2 * n
    assert np.max(np.abs(got - ref)) <= tol
    # structure of the result: right half is exactly the pairing image of the left half
    a, b = got[:m, :m], got[m:, :m]
    assert np.array_equal(got[:m, m:], -b.conj()) and np.array_equal(got[m:, m:], a.conj())
    # the congruence result goes through the solver; eigenvectors are multiplied back: C = X C'
    if m == n:
        import zquatev_b200 as z
        Fp = 0.5 * (ref + ref.conj().T)
        buf = torch.from_numpy(np.ascontiguousarray(Fp.T)).cuda()
        eig = torch.zeros(m, dtype=torch.float64, device="cuda")
        assert z.zquatev_device(2 * m, buf.data_ptr(), 2 * m, eig.data_ptr()) == 0
        Cd = torch.zeros((m, 2 * n), dtype=torch.complex128, device="cuda")
        rc = L.zquatev_b200_qgemm(0, 0, n, m, m, 1.0, Xd.data_ptr(), 2 * n, buf.data_ptr(), 2 * m, 0.0, Cd.data_ptr(), 2 * n, None)
        assert rc == 0
        torch.cuda.synchronize()
        Cp = buf.cpu().numpy().T[:, :m]                   # left half of the eigenvectors C'
        C = Cd.cpu().numpy().T                            # 2n x m
        assert np.max(np.abs(C - X @ np.vstack([Cp[:m], Cp[m:]]))) <= 1e-12 * np.linalg.norm(X, 2) * n


def test_qgemm_argument_checks():
    from zquatev_b200 import api
    L = api.lib()
    assert L.zquatev_b200_qgemm(0, 0, -1, 1, 1, 1.0, None, 2, None, 2, 0.0, None, 2, None) < 0
    assert L.zquatev_b200_congruence(4, 4, None, 8, None, 8, None, 8, None, None) < 0
    assert L.zquatev_b200_fill_pairing(4, 4, None, 8, None) < 0
