"""The reference operator in ONE call across the C ABI: ``sgp_filter`` / ``sgp_filter_host`` (include/sgp_lattice.h),
the C form of ``filter(src, ref, coeffs)`` (gpytorch_lattice_kernel/cpp/lattice.cpp:6-16, cuda/permutohedral_cuda.cpp:12-22).
These tests call nothing else of the library (PyTorch only provides the device buffers) and compare with the oracle's
restatement of the reference filter."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import MAT15_2, MAT15_3, RBF1, make_inputs

pytestmark = pytest.mark.gpu
TOL = 1e-5   # north star: MVM outputs within 1e-5 relative of the reference's filter (fp32)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _vp(t):
    return C.c_void_p(t.data_ptr())


def _rel(got, want):
    return float(np.linalg.norm(got.astype(np.float64) - want) / max(np.linalg.norm(want.astype(np.float64)), 1e-30))


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("N,d,L,coeffs,pad", [
    (2000, 8, 16, RBF1, 0), (777, 3, 1, RBF1, 0), (1500, 11, 11, MAT15_2, 3), (300, 24, 4, MAT15_3, 1),
    (64, 1, 2, [0.5, 1.0, 0.5], 0), (500, 5, 7, [1.0], 2),
])
def test_sgp_filter_one_call_matches_oracle(sg, oracle, N, d, L, coeffs, pad):
    from simplex_gp_b200 import _capi
    lib = _capi.lib()
    x, v = make_inputs(N, d, L, seed=N + d + L)
    want = oracle.filter(v.numpy(), x.numpy(), np.asarray(coeffs, np.float32))
    # padded leading dimensions: the entry takes lds / ldx / ldo like any BLAS-style C interface
    xd = torch.zeros(N, d + pad, device="cuda"); xd[:, :d] = x.cuda()
    vd = torch.zeros(N, L + pad, device="cuda"); vd[:, :L] = v.cuda()
    out = torch.full((N, L + pad), float("nan"), device="cuda")
    c = np.asarray(coeffs, np.float32)
    r = c.shape[0] // 2
    nbytes = lib.sgp_filter_workspace_bytes(N, d, L, r, 0)
    assert nbytes > 0
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    M = C.c_int64(0)
    rc = lib.sgp_filter(_vp(vd), L + pad, _vp(xd), d + pad, _fp(c), c.shape[0], N, L, d, _vp(out), L + pad, _vp(ws), nbytes, 0,
                        C.byref(M), _stream())
    assert rc == 0, lib.sgp_last_error()
    torch.cuda.synchronize()
    assert M.value == oracle.OracleLattice(x.numpy(), coeffs).M
    got = out[:, :L].cpu().numpy()
    assert np.isfinite(got).all()
    assert _rel(got, want) < TOL
    if pad:
        assert torch.isnan(out[:, L:]).all()   # the padding columns are never written


def test_sgp_filter_host_pointers(sg, oracle):
    from simplex_gp_b200 import _capi
    lib = _capi.lib()
    N, d, L = 3000, 8, 16
    x, v = make_inputs(N, d, L, seed=5)
    want = oracle.filter(v.numpy(), x.numpy(), np.asarray(RBF1, np.float32))
    c = np.asarray(RBF1, np.float32)
    for pinned in (False, True):
        xs, vs = (x.pin_memory(), v.pin_memory()) if pinned else (x, v)
        out = torch.empty(N, L, pin_memory=pinned)
        nbytes = lib.sgp_filter_host_workspace_bytes(N, d, L, 1, 0)
        ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        M = C.c_int64(0)
        rc = lib.sgp_filter_host(_vp(vs), L, _vp(xs), d, _fp(c), 3, N, L, d, _vp(out), L, _vp(ws), nbytes, 0, C.byref(M),
                                 _stream())
        assert rc == 0, lib.sgp_last_error()
        assert M.value > 0
        assert _rel(out.numpy(), want) < TOL   # the call returns with the result in host memory


def test_sgp_filter_reports_lattice_size_when_workspace_is_small(sg, oracle):
    from simplex_gp_b200 import _capi
    lib = _capi.lib()
    N, d, L = 4000, 6, 4
    x, v = make_inputs(N, d, L, seed=9)
    xd, vd = x.cuda(), v.cuda()
    out = torch.empty(N, L, device="cuda")
    c = np.asarray(RBF1, np.float32)
    small = lib.sgp_filter_workspace_bytes(N, d, L, 1, 16)
    assert small < lib.sgp_filter_workspace_bytes(N, d, L, 1, 0)
    ws = torch.empty(small, dtype=torch.uint8, device="cuda")
    M = C.c_int64(0)
    rc = lib.sgp_filter(_vp(vd), L, _vp(xd), d, _fp(c), 3, N, L, d, _vp(out), L, _vp(ws), small, 16, C.byref(M), _stream())
    assert rc == _capi.SGP_ENOMEM and b"workspace" in lib.sgp_last_error()
    true_M = oracle.OracleLattice(x.numpy(), RBF1).M
    assert M.value == true_M
    exact = lib.sgp_filter_workspace_bytes(N, d, L, 1, true_M)
    ws = torch.empty(exact, dtype=torch.uint8, device="cuda")
    rc = lib.sgp_filter(_vp(vd), L, _vp(xd), d, _fp(c), 3, N, L, d, _vp(out), L, _vp(ws), exact, true_M, C.byref(M), _stream())
    assert rc == 0, lib.sgp_last_error()
    assert _rel(out.cpu().numpy(), oracle.filter(v.numpy(), x.numpy(), c)) < TOL
    # a workspace smaller than its own layout is refused before any launch
    rc = lib.sgp_filter(_vp(vd), L, _vp(xd), d, _fp(c), 3, N, L, d, _vp(out), L, _vp(ws), exact - 1, true_M, C.byref(M), _stream())
    assert rc == -1


def test_python_filter_goes_through_the_one_call_entry(sg, oracle):
    """``simplex_gp_b200.filter`` keeps the reference's signature; CUDA and CPU tensors, fp32 and fp64, strided inputs."""
    N, d, L = 2500, 8, 5
    x, v = make_inputs(N, d, L, seed=21)
    c = torch.tensor(RBF1)
    want = oracle.filter(v.numpy(), x.numpy(), c.numpy())
    got = sg.filter(v.cuda(), x.cuda(), c)
    assert got.is_cuda and got.dtype == torch.float32 and _rel(got.cpu().numpy(), want) < TOL
    got = sg.filter(v, x, c)
    assert not got.is_cuda and _rel(got.numpy(), want) < TOL
    got = sg.filter(v.double().cuda(), x.double().cuda(), c)
    assert got.dtype == torch.float64 and _rel(got.cpu().numpy(), want) < TOL   # fp32 arithmetic, caller's dtype back
    wide = torch.randn(N, 2 * L).cuda()
    wide[:, ::2] = v.cuda()
    got = sg.filter(wide[:, ::2], x.cuda(), c)   # non-unit column stride
    assert _rel(got.cpu().numpy(), want) < TOL
    # the second call of a shape sizes its workspace from the first one's lattice
    from simplex_gp_b200 import lattice as L_
    assert L_._M_HINT[(N, d, 1)] == oracle.OracleLattice(x.numpy(), RBF1).M


def test_two_devices_in_one_process(sg, oracle):
    """The opt-in for more than 48 KB of dynamic shared memory is per device: a lattice on a second GPU of the same
    process must launch the blur-group / ring kernels as well as on the first (csrc/sgp_groups.cu)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs in one process")
    N, d, L = 20000, 8, 16
    x, v = make_inputs(N, d, L, seed=3)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        lat = sg.Lattice(x.to(dev), RBF1)
        assert lat.groups is not None
        outs.append(lat.mvm(v.to(dev)).cpu())
        outs.append(sg.filter(v.to(dev), x.to(dev), torch.tensor(RBF1)).cpu())
    want = oracle.OracleLattice(x.numpy(), RBF1).mvm(v.numpy())
    for o in outs:
        assert _rel(o.numpy(), want) < TOL
