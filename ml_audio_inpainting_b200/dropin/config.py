"""Drop-in ``config.py``: the reference's constants, unchanged (reference config.py:1-36).

Dataset roots can be overridden with the environment variables ``LIBRISPEECH_ROOT`` /
``LIBRISPEECH_ROOT_PROCESSED`` (the reference hard-codes a Windows path for USER == "Jacob").
"""
import os
from pathlib import Path

PROJECT_ROOT = Path(os.path.dirname(os.path.abspath(__file__)))

USER = os.environ.get("AIP_USER", "")

if USER == "Jacob":
    LIBRISPEECH_ROOT = Path("C:\\Users\\Jacob\\Documents\\2024\\Northeastern\\CS_6140\\Project\\LibriSpeech\\train-clean-100")
    LIBRISPEECH_ROOT_PROCESSED = Path("C:\\Users\\Jacob\\Documents\\2024\\Northeastern\\CS_6140\\Project\\LibriSpeech_PROCESSED\\train-clean-100")
else:
    LIBRISPEECH_ROOT = Path(os.environ.get("LIBRISPEECH_ROOT", "/LibriSpeech/train-clean-100"))
    LIBRISPEECH_ROOT_PROCESSED = Path(os.environ.get("LIBRISPEECH_ROOT_PROCESSED", "/LibriSpeech_PROCESSED/train-clean-100"))

SAMPLE_AUDIO_DIR = LIBRISPEECH_ROOT / "200/126784"
SAMPLE_AUDIO_FILE = SAMPLE_AUDIO_DIR / "200-126784-0006.flac"

OUTPUT_DIR = Path(os.environ.get("AIP_OUTPUT_DIR", str(PROJECT_ROOT / "output")))
os.makedirs(OUTPUT_DIR, exist_ok=True)      # side effect on import, like the reference (config.py:23-24)

DEFAULT_SAMPLE_RATE = 16000
DEFAULT_N_FFT = 512
DEFAULT_HANN_WINDOW_SIZE = 384
DEFAULT_HANN_HOP_LENGTH = 192

DEFAULT_GAP_START_TIME = 2.0
DEFAULT_GAP_DURATION = 0.5

SUPPORTED_FORMATS = [".flac", ".wav", ".mp3"]
