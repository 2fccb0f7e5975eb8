"""CPU restatement of the google/deluca batched-rollout hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``deluca_b200`` imports this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may.  The product path is the CUDA library behind
``include/deluca_b200.h`` and fails loudly when that library is missing.

PARITY PINNING.  The reference (``/root/reference``, deluca v0.0.17) is pure Python on
JAX; JAX/flax are not installable in this image, and the GPC/BPC/LQR/lung modules do not
import even with JAX.  The reference ships golden values only for
``BreathWaveform`` / ``Controller.decay`` / ``proper_time`` / ``LungEnv.time``
(``deluca/tests/lung/core_test.py:43-127``); those are pinned in
``tests/test_oracle_lung.py``.  Every other function here (LDS, GPC, BPC, LQR, PID,
Pendulum, Cartpole, BalloonLung, DelayLung, igpc) is **parity unpinned**: it follows the
reference source line by line (citations in each docstring), and is cross-checked by
invariants (hand adjoint == autograd == finite differences, M=0 GPC == LQR, Riccati
fixed point == scipy DARE), but no reference-produced number exists for it.

Arithmetic dtype is a parameter everywhere: ``np.float32`` mirrors default JAX,
``np.float64`` is the "exact" arm used to show that the 1e-4 tolerance is meaningful.
All functions are batched over a leading env axis ``N`` (the reference's ``jax.vmap``
axis, ``deluca/training/apg.py:91``).
"""
