import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement (oracle/pointops_ref.c). Test infrastructure only."""
    from oracle import pointops_ref
    pointops_ref.build()
    return pointops_ref


@pytest.fixture(scope="session")
def ref_ext():
    """The reference's own CUDA extension compiled for sm_100a (oracle/_ref), or None."""
    import torch
    if not torch.cuda.is_available():
        return None
    from oracle import build_ref_ext
    try:
        return build_ref_ext.load_prebuilt()
    except Exception:  # pragma: no cover
        return None


@pytest.fixture(autouse=True)
def _reset_process_wide_native_state():
    """A test that dies between registering and clearing a process-wide hook (device dropout-seed counter, bf16 weight
    shadows, direct gradient accumulation) must not change what the following tests compute."""
    yield
    import sys
    ops = sys.modules.get("sceneverse_b200.ops")
    if ops is not None:
        ops.DIRECT_GRAD[0] = False
        ops.clear_shadows()
    lib = sys.modules.get("sceneverse_b200._lib")
    if lib is not None and "libsvgps.so" in lib._cache:
        lib.gps().sv_dropout_seed_offset(None)
