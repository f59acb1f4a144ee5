"""Drop-in ``pre_process_dataset.py`` (reference pre_process_dataset.py:19-43): walk LIBRISPEECH_ROOT,
zero a random 0.1 s range per file, peak-normalise, write under LIBRISPEECH_ROOT_PROCESSED.

Same per-file ``np.random`` draws, in the same (os.walk) order, as the reference; the numeric work runs on
the device in batches (``ml_audio_inpainting_b200.preprocess.preprocess_tree``) instead of one file at a time.
"""
from config import LIBRISPEECH_ROOT, LIBRISPEECH_ROOT_PROCESSED, SUPPORTED_FORMATS
from ml_audio_inpainting_b200 import preprocess

if __name__ == "__main__":
    preprocess.preprocess_tree(LIBRISPEECH_ROOT, LIBRISPEECH_ROOT_PROCESSED, gap_len=0.1,
                               supported_formats=SUPPORTED_FORMATS)
