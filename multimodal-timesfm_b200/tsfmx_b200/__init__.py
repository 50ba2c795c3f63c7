"""tsfmx_b200 — B200-native drop-in for the TSFMx multimodal forecast hot path.

Mirrors the reference package layout (`tsfmx.tsfm`, `tsfmx.fusion`, `tsfmx.decoder`, ...); every device
stage behind the adapter API is hand-written CUDA for sm_100a reached through a C ABI (`include/tsfmx_b200.h`).
"""

__version__ = "0.1.0"
