import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

EMU_LIB = os.path.join(ROOT, "tests", "_emu", "libtppvof_emu.so")
CSRC = os.path.join(ROOT, "openfoam-tpp_b200", "csrc")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "gpu2: needs two CUDA devices (gpurun --gpus 2); run with -m gpu2")


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def emu_lib():
    """Host emulation of the kernel bodies (tests only): same source, TPP_EMU build."""
    src = [os.path.join(CSRC, f) for f in ("tppvof.cu", "tpp_kernels.h", "tpp_linsolve.h", "tpp_vcycle.h", "tpp_common.h", "tpp_caseio.h")]
    if not os.path.exists(EMU_LIB) or os.path.getmtime(EMU_LIB) < max(os.path.getmtime(s) for s in src):
        subprocess.run(["make", "-C", CSRC, "emu"], check=True, capture_output=True)
    return EMU_LIB


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library on a real GPU; GPU tests fail loudly if it is missing."""
    from openfoam_tpp_b200 import solver

    assert os.path.exists(solver.LIB_PATH), "libtppvof.so not built (run __graft_entry__.build())"
    assert _has_gpu(), "no CUDA device visible"
    return solver.LIB_PATH
