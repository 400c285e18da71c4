"""CPU oracle for the WaveFormer hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain CPU restatement (PyTorch fp32/fp64 on the host, numpy, and a small C file) of the
algorithm the reference executes on the path in SURVEY.md section 8.  It exists to CHECK the CUDA product in
``waveformer_b200``; it is never the thing measured or shipped.

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py``.  Nothing under ``waveformer_b200/`` imports it, and the product raises if its CUDA library is missing
(there is no CPU fallback).

Pinning status
--------------
* Everything that lives in ``/root/reference`` (attention, Block, PatchMerging, CCF_FFN, the decoder, MONAI's
  sliding-window inferer) is pinned: ``scripts/make_golden.py`` runs the UNMODIFIED reference modules (imported from
  ``/root/reference`` through ``oracle/ref_harness.py``) on seeded inputs and stores their outputs under
  ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this oracle against those fixtures.
* The 3D Haar transform itself lives in the third-party package ``ptwt==0.1.9`` (+ ``PyWavelets==1.6.0``), pinned by
  the reference at ``requirements.txt:45,48`` but absent from ``/root/reference`` and from this image.  ``haar.py``
  restates its published algorithm (``conv_transform_3.py``: separable outer-product filters, ``conv3d`` stride 2,
  ``conv_transpose3d`` stride 2).  No reference test, fixture or golden vector exists at that boundary, so for the
  sub-band sign/naming convention of the helper API the status is **parity unpinned**; what IS pinned is (a) the
  reference's own call sites (``wave_helper.py:350``, ``idwt_upsample.py:160``: key names, tuple order, shapes),
  (b) first-principles known answers (orthonormality, Parseval, impulse/Hadamard table), and (c) whole-model logits,
  which are invariant to the sub-band convention when ``hf_refinement=False``.
"""
