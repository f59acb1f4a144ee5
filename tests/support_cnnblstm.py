"""Test support: the reference-shaped CNN-BLSTM used between front- and back-end (tools/cnnblstm_model.py)."""
from tools.cnnblstm_model import StandInBLSTMCNN  # noqa: F401
