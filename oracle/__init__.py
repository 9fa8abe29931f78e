"""Oracle for the DDSP harmonic-plus-noise synthesis path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the baseline legs of ``bench.py``
(``cpu_baseline``, ``--impl reference`` and ``eager_gpu_baseline``: the reference's
algorithm timed on the host cores, resp. run eagerly on the same GPU) may import
it, and only as the checker or as the thing a baseline times.  The product path (``ddsp_svc_b200``) never imports this package and
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
``frontend``       numpy restatement of the caller-side prologue / epilogue
                   (Volume_Extractor, silence mask, upsample x multiply, cross_fade).
``mel``            STFT.get_mel on the same ATen operators + librosa's mel filterbank
                   restated from its published algorithm (librosa is absent here).
``ref_loader``     imports the live reference from /root/reference -- or, on the
                   GPU box, from the unmodified copy staged under baseline/_ref/ by
                   tools/stage_reference.py (git-ignored; checker / baseline only) --
                   with stub modules for its unused third-party imports.

Parity status: the reference ships no tests, golden vectors or known-answer
fixtures for this path, so parity is pinned on outputs of the
reference itself run in this container (the committed goldens + the live
bit-exactness test), not on reference-owned fixtures.
"""
