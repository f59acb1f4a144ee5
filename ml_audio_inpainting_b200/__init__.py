"""B200-native spectrogram front-end / back-end for savage-hacker14/ml-audio-inpainting.

The package is the GPU replacement of the reference's ``utils.py`` STFT / gap / iSTFT path:

  spectral   batched device-resident stft / istft / griffinlim over the C ABI (include/aip_b200.h)
  gaps       host-side, bit-exact gap / frame index arithmetic
  frontend   batched CNNBLSTM / GAN / eval front-ends and the back-end (callers' epilogues fused)
  dropin/    ``utils.py`` / ``config.py`` / ``add_gaps.py`` / ``pre_process_dataset.py`` with the
             reference's signatures (numpy in, numpy out)

Importing the package does not touch CUDA; the first transform call loads ``lib/libaip_b200.so``
and raises if it is missing or the device is not sm_100 -- there is no CPU fallback.
"""
__version__ = "0.1.0"
