"""Drop-in for 20250506_openLAB_tests/Codes/Models/{temporal_vae_model,cnn_model}.py."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..ops import WindowSource
from ._base import TemporalVAEBase, _HandleModule

SEQ_LEN = 200        # Codes/Models/cnn_model.py:5-6
NUM_FEATURES = 4


class VAE(TemporalVAEBase):
    """Codes/Models/temporal_vae_model.py:4-66."""
    _layer_norm = True

    def __init__(self, input_dim=4, latent_dim=16, hidden_dim=128, num_layers=2, dropout=0.3):
        super().__init__(input_dim, latent_dim, hidden_dim, num_layers, dropout)


class CNN(_HandleModule):
    """Codes/Models/cnn_model.py:8-57.  Input (B, 1, 200, 4) -> logits (B, 2) for [SF, E]."""
    _handle_cls = ops.CnnOpenLab
    ARCH = 1             # _lib.CNN_OPENLAB

    def __init__(self, input_channels=1, num_classes=2, dropout_rate=0.4):
        super().__init__()
        self.drop_p = float(dropout_rate)
        if input_channels != 1 or num_classes != 2:
            raise ops.ShmfastError("the openLAB CNN kernels are specialised to input_channels=1, num_classes=2")

        def block(cin, cout, kt, kf, pt, pf):
            return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=(kt, kf), padding=(pt, pf)),
                                 nn.GroupNorm(num_groups=8, num_channels=cout), nn.SiLU(inplace=True))

        self.features = nn.Sequential(
            block(input_channels, 32, 7, 3, 3, 1), nn.MaxPool2d(kernel_size=(2, 1)),
            block(32, 64, 5, 3, 2, 1), nn.MaxPool2d(kernel_size=(2, 1)),
            block(64, 128, 5, 3, 2, 1), nn.MaxPool2d(kernel_size=(2, 1)),
            block(128, 256, 3, 3, 1, 1), nn.AdaptiveAvgPool2d((1, 1)),
        )
        self.classifier = nn.Sequential(nn.Flatten(), nn.Linear(256, 128), nn.SiLU(inplace=True), nn.Dropout(dropout_rate),
                                        nn.Linear(128, num_classes))
        self.apply(self._init_weights)       # cnn_model.py:45-52
        self._init_handle()

    @staticmethod
    def _init_weights(m):
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            nn.init.kaiming_normal_(m.weight, nonlinearity="relu")
            if m.bias is not None:
                nn.init.zeros_(m.bias)

    def forward(self, x):
        if self.training:
            if x.dim() != 4 or tuple(x.shape[1:]) != (1, SEQ_LEN, NUM_FEATURES):
                raise ops.ShmfastError(f"openLAB CNN expects [B,1,200,4], got {tuple(x.shape)}")
            return self._train_forward(x)
        x = self._eval_only(x)
        if x.dim() != 4 or tuple(x.shape[1:]) != (1, SEQ_LEN, NUM_FEATURES):
            raise ops.ShmfastError(f"openLAB CNN expects [B,1,200,4], got {tuple(x.shape)}")
        return self.handle().forward(WindowSource(x.reshape(x.shape[0], SEQ_LEN, NUM_FEATURES), SEQ_LEN))
