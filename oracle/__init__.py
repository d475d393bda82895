"""CPU oracle for the feature front-end -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker or as
the timed CPU baseline.  The product package never imports it and has no CPU fallback.

Parity status
-------------
* STFT path: **pinned** -- the reference's own code path (``torchaudio.transforms.Spectrogram`` ->
  ``torch.stft``, reference wavelet_math.py:47,63) is importable and is used directly as the checker.
* Wavelet-packet and Haar-fingerprint paths: the arithmetic lives in third-party ``ptwt`` / ``pywt``
  (reference requirements.txt:4-5, un-pinned, un-vendored, no wheel in this image, no network), so neither
  library can be run here.  The oracle restates their published algorithm and is pinned at value level on
  numbers that do not come from this repository (tests/test_published_kats_cpu.py; the same vectors go through
  the CUDA kernels in tests/test_published_kats_gpu.py):
    - the PyWavelets documentation's ``pywt.dwt([3, 7, 1, 1, -2, 5, 4, 6], 'db2')`` printout (an asymmetric
      filter: pins tap orientation, the odd convolution phase, the lo/hi assignment and the high-pass sign),
    - its ``WaveletPacket([1..8], 'db1')`` node values (a, d, aa, ad, aaa, aad) and the natural / frequency
      (Gray-code) leaf orders of levels 2 and 3,
    - ``pywt.wavedec`` / ``dwt_coeff_len`` / ``dwt_max_level`` examples, the ptwt README quick-start example,
    - printed filter banks (db1, db2 = closed form of Daubechies' Table 6.1, sym3 dec_lo/dec_hi, coif5 spot taps),
    - 'reflect' = numpy.pad(mode="reflect"), as pywt documents its extension modes.
  The reference's own tests hold shape assertions only (tests/test_transforms.py:36,51,79,98,124,142), which the
  oracle reproduces (tests/test_oracle_cpu.py::test_reference_shape_kats); the three shipped checkpoints are
  known-answer classification tests on the reference's wav fixtures (tests/test_oracle_cpu.py,
  tests/test_pipeline_gpu.py).  What is still NOT available is an output of a live ptwt/pywt on a full-size frame:
  the long filters' tables (sym5, coif4, db8, sym8) are recalled pywt tables certified by their design equations
  (tools/gen_wavelets.py), and coif6 .. coif10 are the exact coiflets of the same family (pywt's own tables for
  those orders may carry table rounding like its coif5 does, ~1e-8 per tap, inside the 1e-5 tolerance).
"""
