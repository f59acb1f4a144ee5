"""Host-side spectrogram plot (reference utils.py:395-478).  Plotting is outside the GPU path; this exists so
that ``utils.visualize_spectrogram`` keeps working where matplotlib is installed."""
from __future__ import annotations

import numpy as np


def visualize_spectrogram(spectrogram, sample_rate=16000, hop_length=512, title="Spectrogram", power=1.0,
                          in_db=False, y_axis="log", x_axis="time", gap_int=None, save_path=None, figsize=(10, 4)):
    import matplotlib
    matplotlib.use("Agg")
    import matplotlib.pyplot as plt
    if power not in (1, 1.0, 2, 2.0):
        raise ValueError("Power must be 1 (energy) or 2 (power)")
    S = np.abs(np.asarray(spectrogram))
    if not in_db:
        ref = S.max() if S.size else 1.0
        amin = 1e-5 if power in (1, 1.0) else 1e-10
        mult = 20.0 if power in (1, 1.0) else 10.0
        S = mult * np.log10(np.maximum(amin, S)) - mult * np.log10(np.maximum(amin, ref))
        S = np.maximum(S, S.max() - 80.0)
    fig, ax = plt.subplots(figsize=figsize)
    t = np.arange(S.shape[1] + 1) * hop_length / sample_rate
    f = np.linspace(0, sample_rate / 2, S.shape[0] + 1)
    img = ax.pcolormesh(t, f, S, shading="flat")
    if y_axis == "log":
        ax.set_yscale("symlog", linthresh=64.0)
    fig.colorbar(img, ax=ax, format="%+2.0f dB")
    if gap_int is not None:
        for x in gap_int:
            ax.axvline(x=x, color="white", linestyle="--")
    ax.set_title(title)
    ax.set_xlabel("Time (s)")
    ax.set_ylabel("Hz")
    fig.tight_layout()
    if save_path:
        fig.savefig(save_path)
    return fig
