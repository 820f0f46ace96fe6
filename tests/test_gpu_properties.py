"""-m gpu: property tests of the full hot path: whatever the size, panel width, leading dimension and
matrix, ts::zquatev's contract must hold -- n ascending eigenvalues equal to LAPACK's within 1e-12 ||A||, exact quaternion
pairing, small residual / orthogonality, only n eigenvalues written, bit-identical reruns."""
import numpy as np
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st, HealthCheck  # noqa: E402

from oracle import zquatev_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu


def _matrix(n, seed, kind):
    if kind == "sym":
        return O.gen_sym(n, seed)
    rng = np.random.default_rng(seed)
    if kind == "clustered":
        lam = np.round(rng.standard_normal(n), 1)           # many exact repeats -> deflation
    elif kind == "graded":
        lam = np.sign(rng.standard_normal(n)) * 10.0 ** rng.uniform(-12, 0, n)
    else:                                                   # "lowrank": most eigenvalues zero
        lam = np.zeros(n)
        lam[: max(1, n // 5)] = rng.standard_normal(max(1, n // 5))
    return O.gen_spectrum(n, lam, seed)


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(n=st.integers(1, 90), seed=st.integers(0, 10 ** 6), nb=st.sampled_from([0, 1, 5, 16, 32, 64]),
       pad=st.sampled_from([0, 0, 2, 10]), kind=st.sampled_from(["sym", "clustered", "graded", "lowrank"]),
       small=st.booleans())
def test_contract_holds(n, seed, nb, pad, kind, small, monkeypatch):
    from tests import gpu_util as G
    monkeypatch.setenv("ZQ_SMALL_N", "256" if small else "0")
    M = _matrix(n, seed, kind)
    eig, out, info = G.solve_host(M, nb=nb, ld2=2 * n + pad)
    assert info == 0
    assert np.all(eig[n:] == -777.0)                         # only n values are written
    w = np.linalg.eigvalsh(M)[0::2]
    nrm = max(np.abs(w).max(), 1e-300)
    assert np.all(np.diff(eig[:n]) >= 0)
    assert np.max(np.abs(eig[:n] - w)) <= 1e-12 * nrm
    res, orth, pair = O.quality(M, out, eig[:n])
    assert pair == 0.0
    assert res < 1.0 and orth < 3.0, (res, orth)
    eig2, out2, _ = G.solve_host(M, nb=nb, ld2=2 * n + pad)
    assert np.array_equal(eig, eig2) and np.array_equal(out, out2)
