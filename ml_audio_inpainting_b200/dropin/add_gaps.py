"""Drop-in ``add_gaps.py`` (reference add_gaps.py:15-38): zero ``[gap_start, gap_start + gap_duration)``
of a file on the device and write it back un-normalised."""
import numpy as np

from utils import load_audio, _zero_range
from config import SAMPLE_AUDIO_FILE
from ml_audio_inpainting_b200 import audio_io


def insert_gap(audio_path, output_path, gap_start, gap_duration, sample_rate=16000):
    """Insert a gap into a FLAC audio file."""
    print("Loading audio...")
    y, orig_sr = load_audio(audio_path, sample_rate)
    gap_start_idx = int(gap_start * sample_rate)                # add_gaps.py:24
    gap_length = int(gap_duration * sample_rate)                # add_gaps.py:25
    print("Adding gap...")
    head = y[:gap_start_idx]
    tail = y[gap_start_idx + gap_length:]
    # np.concatenate([head, zeros(gap_length), tail]) of the reference: a gap running past the end GROWS the file
    y_new = np.zeros(len(head) + gap_length + len(tail), dtype=np.float64)
    y_new[:len(y)] = _zero_range(y, gap_start_idx, gap_length)[:len(y_new)]
    print("Writing output file...")
    audio_io.write_audio(output_path, y_new, sample_rate, "flac")
    print(f"Processed file saved to {output_path}")


if __name__ == "__main__":
    insert_gap(SAMPLE_AUDIO_FILE, "output/200-126784-0006_W_GAP.flac", 2.0, 5.0)
