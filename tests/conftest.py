import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    """Is there a CUDA device?  Probed WITHOUT the product library (nvidia-smi, then torch), so that a library
    that is missing or fails to load on a GPU box makes the -m gpu tests FAIL at import instead of being
    skipped as "no device"."""
    import shutil
    import subprocess
    smi = shutil.which("nvidia-smi")
    if smi:
        try:
            r = subprocess.run([smi, "-L"], capture_output=True, text=True, timeout=60)
            if r.returncode == 0 and "GPU " in r.stdout:
                return True
        except Exception:
            pass
    try:
        import torch
        return bool(torch.cuda.is_available())
    except ImportError:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
