"""CPU-only checks: the C ABI library loads and exports every symbol of include/deluca_b200.h, the
ctypes layouts match, host-side packing (waveform knots, DARE gains, sharding) is right, and the
world_size-2 statistics reduction works over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from deluca_b200 import _abi
    lib = _abi.lib()   # also verifies the struct layouts against dlc_sizeof()
    hdr = open(os.path.join(ROOT, "include", "deluca_b200.h")).read()
    declared = set(re.findall(r"\b(dlc_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_abi.SYMBOLS), declared ^ set(_abi.SYMBOLS)
    assert lib.dlc_abi_version() == 2
    out = subprocess.run(["nm", "-D", "--defined-only", _abi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (dlc_[a-z0-9_]+)", out))
    assert declared <= exported


def test_argument_errors_need_no_gpu():
    """Argument validation happens before any CUDA call, so it can be exercised on the CPU."""
    from deluca_b200 import _abi
    lib = _abi.lib()
    p = _abi.DlcLdsParams(N=4, T=1, d=0, m=3, traj_stride=1)
    rc = lib.dlc_lds_rollout_f32(ctypes.byref(p), *([None] * 18), None, 0, None)
    assert rc == -2 and b"d and m" in lib.dlc_last_error()
    p = _abi.DlcLdsParams(N=4, T=1, d=2, m=1, traj_stride=0)
    assert lib.dlc_lds_rollout_f32(ctypes.byref(p), *([None] * 18), None, 0, None) == -1
    p = _abi.DlcLdsParams(N=0, T=5, d=2, m=1, traj_stride=1)
    assert lib.dlc_lds_rollout_f32(ctypes.byref(p), *([None] * 18), None, 0, None) == 0   # empty batch
    assert lib.dlc_riccati_backward_f32(1, 3, 5, 5, *([None] * 9), None) == -1
    lp = _abi.DlcLungParams(N=1, T=1, sim=7, nk=5)
    assert lib.dlc_lung_pid_rollout_f32(ctypes.byref(lp), ctypes.byref(_abi.DlcLungBuffers()), None) == -1


def test_fused_wrappers_refuse_cpu_tensors():
    from deluca_b200.fused import lds_rollout
    z = torch.zeros
    with pytest.raises(TypeError, match="CUDA tensor"):
        lds_rollout(z(1, 2, 2), z(1, 2, 1), z(1, 1, 2), z(1, 2), 1, policy="lqr")


def test_product_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "deluca_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, re.M):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_waveform_knots_match_the_oracle_table():
    from deluca_b200.lung.core import BreathWaveform
    from oracle import lung as OL
    pip = np.array([10.0, 35.0, 20.0], np.float32)
    kx, kf, period, xp_in = BreathWaveform.create(pip=torch.tensor(pip), peep=5.0).knots(device="cpu")
    okx, okf, operiod = OL.waveform_knots(pip, 5.0)
    assert np.array_equal(np.array(kx, np.float32), okx) and np.array_equal(kf.numpy(), okf)
    assert np.float32(period) == operiod and xp_in == 1.0
    w = BreathWaveform.create()
    # deluca/tests/lung/core_test.py:55-62
    assert [w.at(t) for t in (0.0, 0.5, 1.0, 1.25, 1.5, 3.0)] == [35.0, 35.0, 35.0, 20.0, 5.0, 35.0]


def test_dare_gain_matches_scipy():
    from deluca_b200.riccati import dare_gain
    from oracle import lds as O
    rng = np.random.default_rng(0)
    A, B, _ = O.lds_default_params(rng, 8, 10, 3, np.float64)
    K = O.lqr_gain(A, B, dtype=np.float64)
    K2, _ = dare_gain(torch.tensor(A), torch.tensor(B))
    np.testing.assert_allclose(K2.numpy(), K, rtol=1e-10, atol=1e-12)


def test_torch_cpu_baseline_matches_numpy_oracle():
    from oracle import lds as O
    from oracle.lds_torch import gpc_rollout_torch
    rng = np.random.default_rng(0)
    N, d, m, T = 16, 10, 3, 40
    A, B, x0 = O.lds_default_params(rng, N, d, m)
    K = O.lqr_gain(A, B)
    W = (0.5 * rng.standard_normal((T, N, d))).astype(np.float32)
    r = O.rollout_lds(A, B, K, x0, W, "gpc", H=3, HH=10, lr_scale=1e-4)
    c, x = gpc_rollout_torch(*(torch.tensor(a) for a in (A, B, K, x0, W)), H=3, HH=10, lr_scale=1e-4)
    np.testing.assert_allclose(c.numpy(), r["costs"], rtol=1e-4, atol=1e-5)


def test_shard_ranges_partition():
    from deluca_b200.sharding import shard_range
    for N, world in ((65536, 8), (10, 3), (7, 8), (0, 2)):
        spans = [shard_range(N, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == N
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from deluca_b200.sharding import shard_range, local_stats, reduce_stats, summarize
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
N, T = 1001, 7
g = torch.Generator().manual_seed(0)
cost_sum = torch.rand(N, generator=g, dtype=torch.float64) * 10   # what every rank would compute globally
status = (torch.arange(N) % 97 == 0).int()
lo, hi = shard_range(N, rank, world)
s = reduce_stats(local_stats(cost_sum[lo:hi], status[lo:hi]))
ref = local_stats(cost_sum, status)
assert torch.allclose(s, ref, rtol=1e-12), (s, ref)
out = summarize(s, T)
assert out["finite_envs"] + out["nonfinite_envs"] == N
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_two_rank_statistics_reduction_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script), ROOT],
                       capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def test_header_is_plain_c_and_links(tmp_path):
    """include/deluca_b200.h must be consumable by a C compiler (no C++ / torch types in any signature) and a C
    program must link against the shared library and get sane answers from the calls that need no GPU."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "abi_check.c"
    src.write_text(
        '#include <stdio.h>\n#include <string.h>\n#include "deluca_b200.h"\n'
        "int main(void) {\n"
        "  DlcLdsParams p; memset(&p, 0, sizeof p);\n"
        "  if (dlc_abi_version() != DLC_ABI_VERSION) return 1;\n"
        "  if (dlc_sizeof(0) != sizeof(DlcLdsParams) || dlc_sizeof(5) != sizeof(DlcOfParams)) return 2;\n"
        "  p.N = -1;\n"
        "  if (dlc_lds_rollout_f32(&p, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0) >= 0) return 3;\n"
        "  if (strlen(dlc_last_error()) == 0) return 4;\n"
        "  if (dlc_gae_f32(4, 0, 0, 0, 0, 0, 0, 0.99f, 0.95f, 0, 0, 0, 0) >= 0) return 5;\n"
        '  printf("ok\\n");\n  return 0;\n}\n')
    exe = tmp_path / "abi_check"
    libdir = os.path.join(ROOT, "deluca_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-ldeluca_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", (out.returncode, out.stdout, out.stderr)


def test_xla_ffi_translation_unit_type_checks():
    """deluca_b200/csrc/xla_ffi/deluca_xla_ffi.cc (the XLA FFI custom-call handlers north_star names) compiles both
    without jaxlib's header (guarded: an empty TU) and against an API-shaped stand-in of `xla/ffi/api/ffi.h`
    (tests/xla_ffi_stub) whose Bind().To() static_asserts that every handler is invocable with exactly the bound
    context / argument / attribute / result types -- i.e. the binding chains and the C-ABI calls are type-checked."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    tu = os.path.join(ROOT, "deluca_b200", "csrc", "xla_ffi", "deluca_xla_ffi.cc")
    base = [gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "include")]
    subprocess.run(base + [tu], check=True)
    r = subprocess.run(base + ["-I" + os.path.join(ROOT, "tests", "xla_ffi_stub"), tu], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # a handler whose signature drifts from its binding must fail the check (the stub is not vacuous)
    src = open(tu).read().replace("int32_t policy, int32_t T, int32_t H,", "int32_t policy, int64_t T, float H,", 1)
    bad = os.path.join(ROOT, "tests", "xla_ffi_stub", "_drifted.cc")
    try:
        open(bad, "w").write(src)
        r = subprocess.run(base + ["-I" + os.path.join(ROOT, "tests", "xla_ffi_stub"), bad], capture_output=True, text=True)
        # (int32 attributes convert implicitly to int64 / float in C++, so invocability alone cannot see THIS drift;
        # what it must see is an arity / buffer-type drift)
        src2 = open(tu).read().replace(".Attr<int32_t>(\"policy\")", "", 1)
        open(bad, "w").write(src2)
        r2 = subprocess.run(base + ["-I" + os.path.join(ROOT, "tests", "xla_ffi_stub"), bad], capture_output=True, text=True)
        assert r2.returncode != 0 and "handler signature does not match" in r2.stderr
    finally:
        if os.path.exists(bad):
            os.remove(bad)
