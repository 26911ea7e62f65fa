"""Drop-in for 4DOF/Scripts/Models/{temporal_vae,cnn_model}.py (same names, ctor signatures, state_dict)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ._base import TemporalVAEBase, _HandleModule

SEQ_LEN = 100        # 4DOF/Scripts/Models/cnn_model.py:4-5
NUM_FEATURES = 12


class TemporalVAE(TemporalVAEBase):
    """4DOF/Scripts/Models/temporal_vae.py:8-77.  Input/output [B, T, D]."""
    _layer_norm = True

    def __init__(self, input_dim: int = 12, latent_dim: int = 16, hidden_dim: int = 128, num_layers: int = 2,
                 dropout: float = 0.3) -> None:
        super().__init__(input_dim, latent_dim, hidden_dim, num_layers, dropout)


VAE = TemporalVAE      # temporal_vae.py:81
__all__ = ["TemporalVAE", "VAE", "CNN", "CNNClassifier", "SEQ_LEN", "NUM_FEATURES"]


class CNN(_HandleModule):
    """4DOF/Scripts/Models/cnn_model.py:8-51.  Input (B, 2, 100, 12) -> logits (B, 2)."""
    _handle_cls = ops.Cnn4dof
    ARCH = 0             # _lib.CNN_4DOF

    def __init__(self, input_channels: int = 2, num_classes: int = 2, dropout_rate: float = 0.5):
        super().__init__()
        self.drop_p = float(dropout_rate)
        if input_channels != 2 or num_classes != 2:
            raise ops.ShmfastError("the 4DOF CNN kernels are specialised to input_channels=2, num_classes=2")
        self.conv1 = nn.Sequential(nn.Conv2d(input_channels, 16, kernel_size=3, padding=1), nn.BatchNorm2d(16), nn.ReLU(),
                                   nn.MaxPool2d(kernel_size=2))
        self.conv2 = nn.Sequential(nn.Conv2d(16, 32, kernel_size=3, padding=1), nn.BatchNorm2d(32), nn.ReLU(),
                                   nn.MaxPool2d(kernel_size=2))
        self.flatten = nn.Flatten()
        self.fc1 = nn.Sequential(nn.Linear(32 * 25 * 3, 128), nn.ReLU(), nn.Dropout(dropout_rate))
        self.fc2 = nn.Linear(128, num_classes)
        self.apply(self._init_weights)       # cnn_model.py:36-43
        self._init_handle()

    def _init_weights(self, m):
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.zeros_(m.bias)

    def _bn_buffers(self):
        return [self.conv1[1].running_mean, self.conv1[1].running_var, self.conv2[1].running_mean, self.conv2[1].running_var]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training:
            if x.dim() != 4 or tuple(x.shape[1:]) != (2, SEQ_LEN, NUM_FEATURES):
                raise ops.ShmfastError(f"4DOF CNN expects [B,2,100,12], got {tuple(x.shape)}")
            return self._train_forward(x)
        return self.handle().forward(self._eval_only(x))


class CNNClassifier(CNN):    # cnn_model.py:55-57
    def __init__(self, dropout_rate: float = 0.5):
        super().__init__(input_channels=2, num_classes=2, dropout_rate=dropout_rate)
