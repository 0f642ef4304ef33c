"""oracle/ -- TEST INFRASTRUCTURE ONLY.  Not part of the product.

CPU (torch-float32 / numpy) restatement of the TACO ``fpv_asymmetry`` hot path
(reference: yinzikang/taco, ``IsaacGymEnvs/isaacgymenvs/tasks/fpv_asymmetry.py`` and
``tasks/control/*.py``).  Every function cites the reference file:line it follows.

Who may import this package: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- as the *checker* or as
the *CPU baseline*, never as something shipped.  ``taco_b200/`` must not import it.

Parity status
-------------
* Control / reward / quaternion leaf functions: PINNED against the reference's own
  torch modules, executed in the build container through import stubs
  (``oracle/ref_loader.py``); the vectors live in ``tests/golden/*.npz`` and the
  generating script is ``oracle/make_golden.py``.
* ``FpvBase`` glue (delay buffer, observation history, resets, commands, mix groups): PINNED against the
  reference's own classes.  ``oracle/fake_gym.py`` is a fake ``isaacgym`` (gymapi / gymtorch) + ``gym.spaces`` over
  which FpvPos / FpvRotate / FpvFlip / FpvMix and VecTask run unmodified on the CPU (``simulate`` = our integrator,
  below); ``oracle/ref_draws.py`` feeds their torch.rand / torch.normal call sites from the shared Philox slot table;
  ``oracle/make_golden_glue.py`` records 72-step trajectories into ``tests/golden/glue_<task>_<mode>.npz``.
  ``tests/test_oracle_glue.py`` holds ``RefFpvEnv(reference_exact=True)`` to them: every integer / mask / index
  exact, floats <= 2e-6 on the first steps and <= 7e-5 over the run.
* Rigid-body step: PhysX is closed source and absent => our own documented integrator
  (``oracle/rigid_body.py``), **parity unpinned** against PhysX by construction.
* Random numbers: the reference draws from torch's global generator (irreproducible
  across implementations); oracle and CUDA kernel both draw from the counter-based
  Philox4x32-10 stream defined in ``oracle/philox.py`` (distributional parity with the
  reference, bit parity between oracle and kernel).
"""
