"""CPU restatement of the google/deluca batched-rollout hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``deluca_b200`` imports this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may.  The product path is the CUDA library behind
``include/deluca_b200.h`` and fails loudly when that library is missing.

PARITY IS PINNED to outputs of the reference itself, run in the build container.  The
reference (``/root/reference``, deluca v0.0.17) is pure Python on JAX; JAX / flax are not
installable in this image and several package ``__init__``s do not import even with JAX
(SURVEY F2/F3), so the reference's OWN SOURCE FILES are executed on ``tests/refshim`` (an
eager, torch-backed stand-in for the slice of jax / flax they use, imported by path) by
``tests/golden/make_ref_fixtures.py``, which writes ``tests/golden/ref_*.npz``:
LDS / GPC / BPC / LQR (``ref_lds``), BalloonLung / DelayLung / lung PID / Expiratory /
``run_controller_scan`` / generic PID (``ref_lung``, ``ref_lung_steps``), ``train_controller``'s
``rollout`` and ``jax.value_and_grad(rollout_parallel)`` (``ref_lung_train``), the classic envs
(``ref_classic``), ``deluca/igpc`` (``ref_igpc``), GRC / SFC / DSC / DRC / LQG (``ref_of``),
``apg_rollout`` / ``apg_loss`` / its gradient and ``gae_advantages`` (``ref_drivers``).
``tests/test_ref_fixtures_oracle.py`` holds every module here to those files (float64: to
rounding, 1e-8 .. 1e-13; the reference's float32 run is compared as well), and
``tests/test_ref_fixtures_gpu.py`` holds the CUDA kernels to them.  The reference's own
known answers (``deluca/tests/lung/core_test.py:43-127``: ``BreathWaveform`` /
``Controller.decay`` / ``proper_time`` / ``LungEnv.time``) are pinned in
``tests/test_oracle_lung.py``.  Besides that, invariants: hand adjoint == autograd == finite
differences, M = 0 GPC == LQR, Riccati fixed point == scipy DARE.

Arithmetic dtype is a parameter everywhere: ``np.float32`` mirrors default JAX,
``np.float64`` is the "exact" arm used to show that the 1e-4 tolerance is meaningful.
All functions are batched over a leading env axis ``N`` (the reference's ``jax.vmap``
axis, ``deluca/training/apg.py:91``).
"""
