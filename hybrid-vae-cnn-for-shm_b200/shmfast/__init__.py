"""shmfast -- B200-native hybrid window scoring (LSTM-VAE gate -> MSE -> threshold -> CNN attribution).

Python host side of libshmfast.so (C ABI in include/shmfast.h): ctypes binding (_lib), operators
(ops), the script-level hot loops (pipeline), window-range sharding (shard), drop-in `Models/`
classes (models.fourdof / models.openlab / models.onedof) and synthetic workloads (synth).
"""
__version__ = "0.1.0"
