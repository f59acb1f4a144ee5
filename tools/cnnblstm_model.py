"""The reference's StackedBLSTMCNN architecture (models/CNNBLSTM/model.py:16-108 built from cnn_blstm.yaml: 3 x conv-BN-ReLU
with 16 / 32 / 64 filters -> 3-layer BiLSTM(257 * 64 = 16 448 -> 128) -> Linear(256 -> 257 * 16) -> conv-BN-ReLU x 2 -> conv),
restated as a plain torch.nn module with the same attribute names, so that a reference state_dict loads into it unchanged
(checked against the reference's own class in tests/test_reference_callers.py).  Random init, stock cuDNN layers: the model is
OUT OF SCOPE; it is only "the thing between" the GPU front-end and back-end in the BASELINE
configs[4] measurement (bench.py leg cnnblstm_e2e) and test.  Not part of the product package."""
import torch
import torch.nn as nn


class StandInBLSTMCNN(nn.Module):
    def __init__(self, freq_bins=257, hidden=128, layers=3, enc=(16, 32), dec=(16, 32)):
        super().__init__()
        self.freq_bins, self.dec0 = freq_bins, dec[0]

        def block(i, o):
            return [nn.Conv2d(i, o, 3, padding=1), nn.BatchNorm2d(o), nn.ReLU()]

        self.encoder = nn.Sequential(*block(1, enc[0]), *block(enc[0], enc[1]), *block(enc[1], hidden // 2))
        self.lstm = nn.LSTM(freq_bins * hidden // 2, hidden, num_layers=layers, batch_first=True, bidirectional=True)
        self.projection = nn.Linear(2 * hidden, freq_bins * dec[0])
        self.decoder = nn.Sequential(*block(dec[0], dec[1]), *block(dec[1], dec[0]), nn.Conv2d(dec[0], 1, 3, padding=1))

    def forward(self, x):                       # [B, 1, F, T]
        b, _, f, t = x.shape
        x = self.encoder(x).permute(0, 3, 1, 2).reshape(b, t, -1)
        x, _ = self.lstm(x)
        x = self.projection(x).view(b, t, self.dec0, f).permute(0, 2, 3, 1)
        return self.decoder(x).squeeze(1)

    def reconstruct_spectrogram(self, log_spectrogram_gap, gap_mask):      # model.py:92-108
        out = self(log_spectrogram_gap.unsqueeze(1))
        gap_mask = gap_mask.float()
        return out * gap_mask + log_spectrogram_gap * (1 - gap_mask)
