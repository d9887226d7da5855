"""CPU oracle for the WarpedGANSpace paired-image training step.

TEST INFRASTRUCTURE ONLY.  Everything under ``oracle/`` is a plain fp32 (or
fp64 on request) torch-on-CPU restatement of the reference algorithm, written
functionally over reference-named state dicts.  It exists so that

  * ``tests/`` can check the CUDA path against it,
  * ``__graft_entry__.smoke()`` can check one small invocation against it,
  * ``bench.py`` can time it as the ``cpu_baseline`` / ``--impl reference`` leg.

Nothing under ``warpedganspace_b200/`` may import it: the product path has no
CPU fallback and raises when the CUDA library is missing.

Parity pinning: the reference ships no tests or golden vectors,
so every oracle function was pinned against the *imported, unmodified
reference modules* in the build container by ``oracle/gen_golden.py``; the
resulting input/output fixtures live in ``tests/golden/`` and are re-checked by
``tests/test_oracle_golden.py`` on every run.  The one exception is StyleGAN2's
two native ops, which have no CPU implementation in the reference: they are
restated from ``models/StyleGAN2/op/upfirdn2d.py:152-186`` (the reference's own
``upfirdn2d_native``) and ``op/fused_bias_act_kernel.cu:25-47``.
"""

from . import support_sets, stylegan2, proggan, sngan, biggan, reconstructor, step  # noqa: F401
