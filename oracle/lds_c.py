"""ctypes loader for oracle/lds_c.c (test infrastructure / CPU baseline; see oracle/__init__.py)."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_build", "liblds_oracle.so")
_lib = None


def _cpu_has_avx2():
    try:
        return " avx2 " in open("/proc/cpuinfo").read()
    except OSError:
        return False


def lib():
    global _lib
    if _lib is None:
        import subprocess
        if not _cpu_has_avx2():   # the prebuilt .so targets x86-64-v3; rebuild generically on older hosts
            subprocess.run(["make", "-C", _HERE, "clean"], check=True, capture_output=True)
            subprocess.run(["make", "-C", _HERE, "MARCH="], check=True, capture_output=True)
        elif not os.path.exists(_PATH):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        _lib = ctypes.CDLL(_PATH)
    return _lib


def max_threads():
    return lib().lds_gpc_max_threads_f32()


def gpc_rollout_c(A, B, K, x0, W, H=3, HH=10, lr_scale=0.005, decay=True, dtype=np.float32, nthreads=0):
    """Fresh-agent LDS+GPC rollout.  Returns costs[T,N], x_final[N,d]."""
    T, N, d = W.shape
    m = B.shape[2]
    c = lambda a: np.ascontiguousarray(a, dtype=dtype)
    A, B, K, W, x = c(A), c(B), c(K), c(W), c(x0).copy()
    costs = np.empty((T, N), dtype=dtype)
    fn = lib().lds_gpc_rollout_f32 if dtype == np.float32 else lib().lds_gpc_rollout_f64
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = fn(ctypes.c_int64(N), T, d, m, H, HH, p(A), p(B), p(K), p(x), p(W), p(costs), ctypes.c_double(lr_scale),
            int(decay), int(nthreads))
    if rc != 0:
        raise ValueError("unsupported shape for the C oracle")
    return costs, x
