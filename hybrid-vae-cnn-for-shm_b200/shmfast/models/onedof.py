"""Drop-in for 1_DOF/Scripts/Models/temporal_vae.py (no LayerNorm, static reparameterize)."""
from __future__ import annotations

from ._base import TemporalVAEBase


class TemporalVAE(TemporalVAEBase):
    """1_DOF/Scripts/Models/temporal_vae.py:8-58."""
    _layer_norm = False

    def __init__(self, input_dim: int = 12, latent_dim: int = 5, hidden_dim: int = 32, num_layers: int = 2,
                 dropout: float = 0.2):
        super().__init__(input_dim, latent_dim, hidden_dim, num_layers, dropout)
