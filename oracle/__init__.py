"""CPU oracle for the spectrogram front-end / back-end (TEST INFRASTRUCTURE ONLY).

This package restates, in numpy + scipy.fft on the CPU, the reference's hot path:
``utils.py`` (gap masks, STFT, iSTFT / phase reuse / Griffin-Lim), the per-item
epilogues of ``models/CNNBLSTM/dataset.py`` / ``models/GAN/dataset.py`` /
``models/model_eval.py`` and the third-party ``librosa`` routines those call
(``librosa_port``, ``utils_port``, ``callers_port``), plus ``flac_port``: the pure-Python FLAC
codec that checks the product's native host codec (``csrc/aip_flac.c``).

Rules (enforced by tests/test_host_logic.py):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
    ``--impl reference`` legs may import anything from here, and only as the checker
    or as the CPU baseline -- never as the thing shipped or measured as "ours";
  * the product package (``ml_audio_inpainting_b200``) must not import it and has no
    CPU fallback: it raises when the CUDA library is missing.

Parity pin status
-----------------
The arithmetic of this path lives in ``librosa`` (``requirements.txt:3``,
``librosa>=0.8.1``, unpinned upper bound; restated here at librosa >= 0.10
semantics), which is NOT vendored under /root/reference and NOT installable in this
image (no network), so librosa itself cannot be executed here.  The oracle is
nevertheless PINNED TO VALUES THE REFERENCE ITSELF PRODUCED:

  * ``/root/reference/test_samples_reconstructed/*_cnnlstm_inpainted.flac`` -- nine files
    written by the reference's ``models/model_eval.py:179-192`` with the real librosa and
    soundfile.  ``reconstruct_spectrogram`` (``models/CNNBLSTM/model.py:108``) keeps its input
    outside the gap frames [166, 173), so away from the gap each file is the reference's own
    load -> STFT -> log10 -> 10** -> phase reuse -> iSTFT -> peak normalise -> PCM-16 of the
    matching ``test_samples`` clip.  The oracle reproduces all nine with NO free parameter
    (the peak of every file lies outside the gap): at most 1 LSB of 16-bit PCM apart, on
    2 ... 54 of 77 264 samples per clip (tests/test_reference_outputs.py; fixture
    tests/golden/reference_cnnlstm_inpainted_int16.npz).  This pins stft, istft, the window,
    the centre padding, time_to_frames, normalize and the FLAC PCM scale (x 32768, clipped).
  * the reference's own CALLER SOURCE, executed unmodified in the build container
    (tests/refshim.py stands in for the three absent third-party packages only):
    ``LibriSpeechDataset.__getitem__``, ``SpeechInpaintingDataset.__getitem__``,
    ``model_eval.inpaint`` (both branches) and the ``pre_process_dataset.py`` loop body
    produce tests/golden/reference_callers.npz; ``callers_port.py`` / ``utils_port.py``
    reproduce it bit for bit (tests/test_reference_callers.py), which is what entitles the
    GPU tests to use these restatements as their checker.
  * the reference's own test-suite (tests/utils_test.py) run unchanged against the
    reference's utils.py on this oracle: 19 of 36 pass; the 17 others fail for reasons
    listed in tests/test_gpu_utils_curated.py (stale against the shipped utils.py, plotting,
    or Griffin-Lim from unseeded random phases on a complex "magnitude").

What remains unpinned by reference-produced values: Griffin-Lim's VALUES (the reference
ships no Griffin-Lim output and draws unseeded random phases) and the mel functions (no
shipped output); for those the pins are independent implementations:
  * ``torch.stft`` / ``torch.istft`` on CPU agree with ``librosa_port.stft`` / ``istft`` to
    ~1e-7 relative max-abs (tests/test_oracle.py);
  * ``torchaudio.functional.melscale_fbanks(norm='slaney', mel_scale='slaney')`` agrees with
    ``librosa_port.mel`` to 6e-6 (tests/test_oracle.py::test_mel_matches_torchaudio);
  * the two tight properties the reference's own tests assert on librosa
    (``tests/utils_test.py:780-809`` and ``:811-849``: float64 round trip with
    atol=1e-10 at n_fft 512 / hop 192 / win 384) hold for the restatement.
"""
