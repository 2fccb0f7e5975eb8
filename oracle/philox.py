"""Counter-based disturbance generator (test infrastructure; see oracle/__init__.py).

The reference draws LDS disturbances from ``jax.random.normal(key, (d, 1)) * std`` with a
key that is split once per step (``deluca/envs/_lds.py:35-41,102-106``).  Threefry streams
cannot be reproduced without JAX, so the CUDA path defines its own
counter-based stream and this module restates it on the CPU:

  Philox4x32-10 (Salmon et al., SC'11; the published round function and constants),
  key     = (seed_lo, seed_hi)
  counter = (env, t, j, 0)          j = b // 4 selects the group of four coordinates
  four 32-bit outputs (r0..r3) -> two Box-Muller pairs:
      u1 = ((r >> 8) + 1) * 2^-24  in (0, 1]     u2 = (r' >> 8) * 2^-24  in [0, 1)
      ang = 2 pi (u2 - 1/2) in [-pi, pi)
      z0 = sqrt(-2 ln u1) cos(ang),  z1 = sqrt(-2 ln u1) sin(ang)
  w[t, env, b] = std * z[b % 4]

The integer part is bit-exact against the kernel; the float transform uses libm here and
MUFU intrinsics on the GPU, so values agree to ~1e-6 absolute (tested), not bitwise.
"""
import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32(counter, key, rounds=10):
    """counter: uint32[..., 4], key: uint32[..., 2] -> uint32[..., 4]."""
    c = [np.asarray(counter[..., i], dtype=np.uint32).copy() for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint32).copy()
    k1 = np.asarray(key[..., 1], dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        for _ in range(rounds):
            p0 = c[0].astype(np.uint64) * PHILOX_M0
            p1 = c[2].astype(np.uint64) * PHILOX_M1
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & _MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & _MASK32).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = (k0 + PHILOX_W0).astype(np.uint32)
            k1 = (k1 + PHILOX_W1).astype(np.uint32)
    return np.stack(c, axis=-1)


def box_muller_f32(r_a, r_b):
    """Two uint32 arrays -> two float32 standard normals (see module docstring)."""
    f32 = np.float32
    u1 = ((r_a >> np.uint32(8)).astype(np.uint32) + np.uint32(1)).astype(f32) * f32(2.0 ** -24)
    u2 = (r_b >> np.uint32(8)).astype(f32) * f32(2.0 ** -24)
    rad = np.sqrt(f32(-2.0) * np.log(u1)).astype(f32)
    ang = (f32(2.0 * np.pi) * (u2 - f32(0.5))).astype(f32)
    return (rad * np.cos(ang)).astype(f32), (rad * np.sin(ang)).astype(f32)


def gaussian_disturbance(seed, env_ids, t0, T, d, std=0.5):
    """W[T, N, d] float32 -- the stream the kernels generate in-register.

    ``env_ids`` are *global* env indices (so a shard of envs draws the same numbers on
    any rank); ``t0`` is the step counter at the start of the rollout.
    """
    env_ids = np.asarray(env_ids, dtype=np.uint32)
    N = env_ids.shape[0]
    ngrp = (d + 3) // 4
    tt = (np.arange(T, dtype=np.uint64) + np.uint64(t0)).astype(np.uint32)
    ctr = np.zeros((T, N, ngrp, 4), dtype=np.uint32)
    ctr[..., 0] = env_ids[None, :, None]
    ctr[..., 1] = tt[:, None, None]
    ctr[..., 2] = np.arange(ngrp, dtype=np.uint32)[None, None, :]
    key = np.zeros((T, N, ngrp, 2), dtype=np.uint32)
    key[..., 0] = np.uint32(seed & 0xFFFFFFFF)
    key[..., 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    r = philox4x32(ctr, key)
    z0, z1 = box_muller_f32(r[..., 0], r[..., 1])
    z2, z3 = box_muller_f32(r[..., 2], r[..., 3])
    z = np.stack([z0, z1, z2, z3], axis=-1).reshape(T, N, ngrp * 4)[..., :d]
    return (np.float32(std) * z).astype(np.float32)
