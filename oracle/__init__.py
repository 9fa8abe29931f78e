"""Oracle for the DDSP harmonic-plus-noise synthesis path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline / reference
arm of ``bench.py`` may import it, and only as the checker or the timed CPU
baseline.  The product path (``ddsp_svc_b200``) never imports this package and
fails loudly when its CUDA library is missing.

Contents
--------
``synth_inputs``   seeded synthetic control tensors.
``torch_port``     restatement of the reference algorithm with the same ATen CPU
                   operators the reference calls (the arithmetic of the path lives
                   in PyTorch, which is not vendored under /root/reference); it is
                   pinned bit-for-bit against the live reference in this container
                   (tests/test_oracle_vs_reference.py) and against committed golden
                   vectors generated from the live reference
                   (tests/golden/, made by tests/golden/make_golden.py).
``closed_form``    independent float64 numpy restatement of the closed-form math
                    -- the tie-breaker / ground truth.
``ref_loader``     imports the live reference from /root/reference with stub
                   modules for its unused third-party imports; only usable in the
                   build container (the reference does not travel to the GPU box).

Parity status: the reference ships no tests, golden vectors or known-answer
fixtures for this path, so parity is pinned on outputs of the
reference itself run in this container (the committed goldens + the live
bit-exactness test), not on reference-owned fixtures.
"""
