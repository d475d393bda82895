"""CPU oracle for the feature front-end -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker or as
the timed CPU baseline.  The product package never imports it and has no CPU fallback.

Parity status
-------------
* STFT path: **pinned** -- the reference's own code path (``torchaudio.transforms.Spectrogram`` ->
  ``torch.stft``, reference wavelet_math.py:47,63) is importable and is used directly as the checker.
* Wavelet-packet and Haar-fingerprint paths: the arithmetic lives in third-party ``ptwt`` / ``pywt``
  (reference requirements.txt:4-5, un-pinned, un-vendored, not installable offline).  The oracle restates
  their published algorithm.  The reference's own tests hold shape assertions only
  (tests/test_transforms.py:36,51,79,98,124,142), which the oracle reproduces; value-level pins are the
  three shipped checkpoints (known-answer classification of the reference's wav fixtures, see
  tests/test_checkpoint_kat.py) and pywt's documented Haar identity.  Beyond those anchors the packet
  values are **parity unpinned** against a live ptwt.
"""
