"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the FRNet hot path.

This package restates, on the CPU in fp32, the algorithm of the reference
generator hot path (skycrapers/TecoGAN-PyTorch @ 903b070):

    FNet -> reflect-pad -> bicubic/bilinear upsample of the flow -> backward_warp
         -> space_to_depth + concat -> SRNet (conv_in, residual blocks,
            stride-2 transposed convs, conv_out, + upsampled LR residual)
         -> float32_to_uint8

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product package
(``tecogan-pytorch_b200``) never imports anything from here, and fails loudly
when its CUDA library is missing.

Parity pin: the reference repo has no tests and no golden vectors
(SURVEY.md section 4).  The oracle is pinned against OUTPUTS OF THE REFERENCE
ITSELF: ``oracle/gen_golden.py`` imports the unmodified reference from
``/root/reference`` in the build container, runs it on seeded inputs and
writes the fixtures in ``tests/golden``; ``tests/test_oracle_golden.py``
checks every oracle function against those fixtures (and, when
``/root/reference`` is present, directly against the live reference at the
BASELINE sizes).
"""
