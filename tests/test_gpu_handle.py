"""-m gpu: the handle / workspace API (the reference allocates per call, zquatev.cc:63-66): plans of several
sizes stay cached side by side, handles do not share state, asynchronous solves on different streams through one
handle are ordered on the device, and the robustness the reference gets from LAPACK's scaling (zlascl inside zheev /
zhbev, scaled dznrm2 / zlarfg at blocked.cc:176,345)."""
import ctypes
import threading

import numpy as np
import pytest

from oracle import zquatev_oracle as O

pytestmark = pytest.mark.gpu


def _solve_handle(h, M, opt=None):
    from zquatev_b200 import api
    n2 = M.shape[0]
    buf = np.asfortranarray(M).copy(order="F")
    eig = np.zeros(n2)
    o = opt or api.ZqOptions(1, 0, 0, None, 1, 0, 0, 0, 0)
    info = api.lib().zquatev_b200_solve(h, n2, buf.ctypes.data, n2, eig.ctypes.data, ctypes.byref(o))
    return eig[: n2 // 2], buf, info


def test_handle_alternating_sizes_and_query():
    import torch
    from zquatev_b200 import api
    L = api.lib()
    h = ctypes.c_void_p()
    assert L.zquatev_b200_create(ctypes.byref(h)) == 0
    sizes = [40, 300, 65]
    Ms = {n: O.gen_sym(n, 50 + n) for n in sizes}
    ref = {n: np.linalg.eigvalsh(Ms[n])[0::2] for n in sizes}
    for n in sizes:                      # first pass allocates the three plans
        e, out, info = _solve_handle(h, Ms[n])
        assert info == 0 and np.max(np.abs(e - ref[n])) <= 1e-12 * np.abs(ref[n]).max()
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(3):                   # alternating sizes: no plan is dropped or re-allocated
        for n in sizes:
            e, out, info = _solve_handle(h, Ms[n])
            assert info == 0 and np.max(np.abs(e - ref[n])) <= 1e-12 * np.abs(ref[n]).max()
            assert O.quality(Ms[n], out, e)[2] == 0.0
    torch.cuda.synchronize()
    assert torch.cuda.mem_get_info()[0] == free0, "alternating sizes must not allocate"
    ms = (ctypes.c_double * 8)()
    assert L.zquatev_b200_handle_phases(h, ms) == 1 and ms[5] > 0.0
    # workspace query: grows with n, host-pointer mode adds the 2n x 2n staging array, values-only drops the D&C
    q = ctypes.c_ulonglong(0)
    od = api.ZqOptions(1, 1, 0, None, 1, 0, 0, 0, 0)
    oh = api.ZqOptions(1, 0, 0, None, 1, 0, 0, 0, 0)
    ov = api.ZqOptions(0, 1, 0, None, 1, 0, 0, 0, 0)
    assert L.zquatev_b200_workspace_query(2 * 300, ctypes.byref(od), ctypes.byref(q)) == 0
    dev300 = q.value
    assert L.zquatev_b200_workspace_query(2 * 300, ctypes.byref(oh), ctypes.byref(q)) == 0
    assert q.value == dev300 + 16 * 600 * 600
    assert L.zquatev_b200_workspace_query(2 * 300, ctypes.byref(ov), ctypes.byref(q)) == 0
    assert 0 < q.value < dev300
    assert L.zquatev_b200_workspace_query(2 * 600, ctypes.byref(od), ctypes.byref(q)) == 0 and q.value > dev300
    assert L.zquatev_b200_reserve(h, 2 * 500, ctypes.byref(oh)) == 0          # allocate ahead of the first solve
    free1 = torch.cuda.mem_get_info()[0]
    e, out, info = _solve_handle(h, O.gen_sym(500, 3))
    torch.cuda.synchronize()
    assert info == 0 and torch.cuda.mem_get_info()[0] == free1
    assert L.zquatev_b200_destroy(h) == 0


def test_two_handles_concurrent_threads():
    """two host threads, one handle each: no shared plan, no process-wide lock -- both must get their own answers"""
    from zquatev_b200 import api
    L = api.lib()
    res = {}

    def work(tag, n, seed):
        import torch
        torch.cuda.set_device(0)
        h = ctypes.c_void_p()
        assert L.zquatev_b200_create(ctypes.byref(h)) == 0
        ok = True
        for it in range(6):
            M = O.gen_sym(n, seed + it)
            e, out, info = _solve_handle(h, M)
            w = np.linalg.eigvalsh(M)[0::2]
            ok = ok and info == 0 and np.max(np.abs(e - w)) <= 1e-12 * np.abs(w).max() and O.quality(M, out, e)[2] == 0.0
        L.zquatev_b200_destroy(h)
        res[tag] = ok

    ts = [threading.Thread(target=work, args=("a", 150, 100)), threading.Thread(target=work, args=("b", 290, 200))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert res == {"a": True, "b": True}


def test_async_solves_on_two_streams_share_a_plan_safely():
    """device-pointer mode, sync = 0: the second solve (same size, other stream) reuses the plan of the first, which
    is still running; it must be ordered behind it on the device (silent race)."""
    import torch
    import zquatev_b200 as z
    n = 400
    Ms = [O.gen_sym(n, 300 + i) for i in range(2)]
    bufs = [torch.from_numpy(np.asfortranarray(M).T.copy()).cuda() for M in Ms]
    eigs = [torch.zeros(n, dtype=torch.float64, device="cuda") for _ in Ms]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    z.zquatev_device(2 * n, bufs[0].data_ptr(), 2 * n, eigs[0].data_ptr(), stream=s1.cuda_stream, sync=False)
    z.zquatev_device(2 * n, bufs[1].data_ptr(), 2 * n, eigs[1].data_ptr(), stream=s2.cuda_stream, sync=False)
    torch.cuda.synchronize()
    for M, b, e in zip(Ms, bufs, eigs):
        w = np.linalg.eigvalsh(M)[0::2]
        eh = e.cpu().numpy()
        assert np.max(np.abs(eh - w)) <= 1e-12 * np.abs(w).max()
        res, orth, pair = O.quality(M, b.cpu().numpy().T, eh)
        assert pair == 0.0 and res < 1.0 and orth < 2.0


@pytest.mark.parametrize("scale", [1e-160, 1e-120, 1e120, 1e155])
def test_tiny_and_huge_norms(scale, reduction_path):
    """entries far outside [1e-146, 1e145]: the plain sums of squares of the reflector kernels would underflow to 0
    (tau = 0, silently wrong tridiagonal) or overflow; the input scaling (scale.cu, zlascl analogue) prevents both.
    Eigenvalues scale exactly with the matrix, eigenvectors are unchanged."""
    from tests import gpu_util as G
    n = 70
    M0 = O.gen_sym(n, 77)
    e0, o0, i0 = G.solve_host(M0)
    e1, o1, i1 = G.solve_host(M0 * scale)
    assert i0 == 0 and i1 == 0
    assert np.all(np.isfinite(e1[:n])) and np.all(np.isfinite(o1))
    assert np.max(np.abs(e1[:n] / scale - e0[:n])) <= 1e-12 * np.abs(e0[:n]).max()
    res, orth, pair = O.quality(M0, o1, e1[:n] / scale)
    assert pair == 0.0 and res < 1.0 and orth < 2.0


def test_graded_spectrum(reduction_path):
    """strongly graded spectrum (15 decades)"""
    from tests import gpu_util as G
    n = 60
    lam = np.concatenate([-np.logspace(0, -15, n // 2), np.logspace(-15, 0, n - n // 2)])
    M = O.gen_spectrum(n, lam, 3)
    eig, out, info = G.solve_host(M)
    w = np.linalg.eigvalsh(M)[0::2]
    assert info == 0 and np.max(np.abs(eig[:n] - w)) <= 1e-12 * np.abs(w).max()
    res, orth, pair = O.quality(M, out, eig[:n])
    assert pair == 0.0 and res < 1.0 and orth < 2.0
