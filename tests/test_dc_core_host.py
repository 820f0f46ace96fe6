"""CPU test of the __host__ __device__ numerical cores of the device D&C (dc_core.cuh), compiled
for the host by tests/host_shim (g++), against LAPACK through the oracle's D&C flow."""
import numpy as np
import pytest

from oracle import tridiag_dc as T
from tests.host_shim.build import HostCores
from tests.test_oracle import _tri_cases


@pytest.fixture(scope="module")
def cores():
    return HostCores()


@pytest.mark.parametrize("name,d,e", list(_tri_cases()))
def test_cores_vs_lapack(cores, name, d, e):
    st = []
    w, Z = T.stedc(d, e, 32, st, cores)
    n = len(d)
    Tm = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    wr = np.linalg.eigvalsh(Tm)
    nrm = max(np.abs(wr).max(), 1e-300)
    assert np.abs(w - wr).max() <= 50 * T.EPS * nrm
    assert np.linalg.norm(Tm @ Z - Z * w) / (n * nrm * T.EPS) < 2.0
    assert np.linalg.norm(Z.T @ Z - np.eye(n)) / (n * T.EPS) < 2.0
    its = [x[1] for x in st if x[0] == "iters"]
    assert all(i < 100 for i in its)


def test_leaf_sizes(cores):
    rng = np.random.default_rng(3)
    for m in [1, 2, 3, 16, 31, 32]:
        d = rng.standard_normal(m)
        e = rng.standard_normal(max(m - 1, 0))
        w, Z = cores.leaf(d, e)
        Tm = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        assert np.allclose(np.sort(w), np.linalg.eigvalsh(Tm), atol=1e-13)
        assert np.allclose(Tm @ Z, Z * w, atol=1e-13)


def test_secular_interlacing(cores):
    rng = np.random.default_rng(4)
    k = 40
    dl = np.sort(rng.standard_normal(k))
    z = rng.standard_normal(k)
    z /= np.linalg.norm(z)
    rho = 0.7
    org, mu, worst = cores.secular(dl, z * z, rho)
    lam = dl[org] + mu
    ref = np.linalg.eigvalsh(np.diag(dl) + rho * np.outer(z, z))
    assert worst < 100
    assert np.allclose(lam, ref, atol=1e-14 * max(1, np.abs(ref).max()))
    assert np.all(lam[:-1] > dl[:-1]) and np.all(lam[:-1] < dl[1:]) and lam[-1] > dl[-1]
