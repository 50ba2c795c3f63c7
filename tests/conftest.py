"""pytest configuration: the `gpu` marker and import paths.

`-m "not gpu"` runs here (CPU only): oracle vs golden vectors, host logic, C-ABI symbol export.
`-m gpu` runs on a B200: parity of the CUDA path against the oracle, through the C ABI.
"""

import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "multimodal-timesfm_b200", ROOT):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
