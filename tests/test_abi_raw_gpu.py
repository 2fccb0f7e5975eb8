"""The C ABI is array-library-agnostic: this test drives `dlc_lds_rollout_f32` with RAW device pointers from
`cudaMalloc` (libcudart through ctypes) on a stream it created itself -- no torch tensor, no DLPack, nothing but
integers crossing the boundary -- and checks the result against the CPU oracle.  A second case hands the library a
foreign `__cuda_array_interface__` producer (a tiny class defined here), which is how a JAX or CuPy array would arrive.
"""
import ctypes
import os

import numpy as np
import pytest

from oracle import lds as O

pytestmark = pytest.mark.gpu


def _cudart():
    for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            return ctypes.CDLL(name)
        except OSError:
            continue
    import torch  # torch bundles a runtime: load that shared object by path (still no torch tensors below)
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    for f in sorted(os.listdir(libdir)):
        if f.startswith("libcudart"):
            return ctypes.CDLL(os.path.join(libdir, f))
    pytest.skip("no libcudart")


class RawBuffer:
    """cudaMalloc'd memory exposing `__cuda_array_interface__` v3 -- a stand-in for any foreign array library."""

    def __init__(self, rt, host):
        self.rt, self.shape, self.dtype = rt, host.shape, host.dtype
        self.nbytes = host.nbytes
        self.ptr = ctypes.c_void_p()
        assert rt.cudaMalloc(ctypes.byref(self.ptr), ctypes.c_size_t(max(self.nbytes, 4))) == 0
        assert rt.cudaMemcpy(self.ptr, host.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(self.nbytes), 1) == 0  # H2D

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": self.dtype.str, "data": (self.ptr.value, False), "version": 3}

    def get(self):
        out = np.empty(self.shape, self.dtype)
        assert self.rt.cudaMemcpy(out.ctypes.data_as(ctypes.c_void_p), self.ptr, ctypes.c_size_t(self.nbytes), 2) == 0
        return out

    def free(self):
        self.rt.cudaFree(self.ptr)


def test_rollout_from_raw_cudamalloc_pointers():
    from deluca_b200 import _abi
    rt = _cudart()
    lib = _abi.lib()
    rng = np.random.default_rng(0)
    N, d, m, T, H, HH = 37, 10, 3, 120, 3, 10
    A, B, x0 = O.lds_default_params(rng, N, d, m)
    K = O.lqr_gain(A, B)
    W = (0.5 * rng.standard_normal((T, N, d))).astype(np.float32)
    x0 = (0.1 * rng.standard_normal((N, d))).astype(np.float32)
    ref = O.rollout_lds(A, B, K, x0, W, "gpc", H=H, HH=HH, lr_scale=1e-4, dtype=np.float64)
    bufs = {k: RawBuffer(rt, np.ascontiguousarray(v)) for k, v in dict(
        A=A, B=B, K=K, W=W, x=x0, M=np.zeros((N, H, m, d), np.float32), hist=np.zeros((N, H + HH, d), np.float32),
        xprev=np.zeros((N, d), np.float32), costs=np.zeros((T, N), np.float32), cost_sum=np.zeros(N, np.float64),
        status=np.zeros(N, np.int32)).items()}
    stream = ctypes.c_void_p()
    assert rt.cudaStreamCreate(ctypes.byref(stream)) == 0
    p = _abi.DlcLdsParams(N=N, T=T, d=d, m=m, H=H, HH=HH, policy=1, decay=1, lr_scale=1e-4, w_std=0.5, traj_stride=1)
    assert lib.dlc_lds_workspace_bytes(ctypes.byref(p)) == 0
    ptr = lambda k: bufs[k].__cuda_array_interface__["data"][0]        # what a JAX / CuPy binding would pass
    rc = lib.dlc_lds_rollout_f32(ctypes.byref(p), ptr("A"), ptr("B"), ptr("K"), ptr("W"), None, ptr("x"), ptr("M"),
                                 ptr("hist"), ptr("xprev"), None, None, None, None, ptr("costs"), ptr("cost_sum"),
                                 None, None, ptr("status"), None, 0, stream)
    _abi.check(rc, "dlc_lds_rollout_f32")
    assert rt.cudaStreamSynchronize(stream) == 0
    costs = bufs["costs"].get()
    rel = np.abs(costs - ref["costs"]) / np.maximum(np.abs(ref["costs"]), 1e-2 * np.abs(ref["costs"]).mean())
    assert np.isfinite(costs).all() and rel.max() < 1e-4
    np.testing.assert_allclose(bufs["x"].get(), ref["x_final"], rtol=1e-4, atol=1e-4 * np.abs(ref["x_final"]).mean())
    np.testing.assert_allclose(bufs["cost_sum"].get(), ref["costs"].sum(0), rtol=1e-4)
    assert (bufs["status"].get() == 0).all()
    # argument errors come back as negative codes with a message, never as exceptions across the ABI
    p_bad = _abi.DlcLdsParams(N=N, T=T, d=0, m=m, H=H, HH=HH, policy=1, traj_stride=1)
    rc = lib.dlc_lds_rollout_f32(ctypes.byref(p_bad), ptr("A"), ptr("B"), ptr("K"), None, None, ptr("x"), ptr("M"),
                                 ptr("hist"), ptr("xprev"), None, None, None, None, None, None, None, None, None,
                                 None, 0, stream)
    assert rc < 0 and b"d and m" in lib.dlc_last_error()
    rt.cudaStreamDestroy(stream)
    for b in bufs.values():
        b.free()
