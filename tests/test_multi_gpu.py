"""Multi-GPU exchange step through the C-ABI: rb200_comm_* / rb200_gather_plan / rb200_gather_blobs over
NCCL, driven by the C++ multi-process driver (csrc/exam_gpu_multi.cpp: one process per GPU, shards on chunk boundaries,
rank 0 decodes the gathered container).  world = 1 runs wherever one GPU is visible; world = 2 needs two."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _driver():
    import ryg_rans_b200 as rb
    rb.build()
    from ryg_rans_b200.build import EXAM_MULTI_PATH
    assert os.path.exists(EXAM_MULTI_PATH), "exam_gpu_multi was not built"
    return EXAM_MULTI_PATH


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.gpu
@pytest.mark.parametrize("world,coder,chunk,total", [(1, "word", 4096, 1_000_003), (1, "alias", 8192, 3_000_017),
                                                     (2, "word", 4096, 5_000_011), (2, "alias", 8192, 5_000_011),
                                                     (2, "word", 8192, 8192 * 3 + 17)])      # rank 1 gets a single ragged chunk
def test_gather_over_nccl_cpp_driver(world, coder, chunk, total):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    out = subprocess.run([_driver(), str(world), coder, str(chunk), str(total)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "decode ok!" in out.stdout
    assert out.stdout.count("[rank ") == world


def test_gather_symbols_are_exported():
    """No GPU needed: the library exports the exchange entry points and reports a missing GPU / NCCL loudly."""
    import ctypes as C
    import numpy as np
    import ryg_rans_b200 as rb
    lib = rb.load()
    for name in ("rb200_comm_unique_id", "rb200_comm_create", "rb200_comm_destroy", "rb200_gather_plan", "rb200_gather_blobs"):
        assert hasattr(lib.dll, name)
    ident = np.zeros(128, np.uint8)
    rc = lib.dll.rb200_comm_unique_id(ident.ctypes.data)
    assert rc in (0, -8)                                   # an id, or "NCCL error" when libnccl.so.2 is absent
    assert lib.dll.rb200_comm_create(None, ident.ctypes.data, 0, 1, C.byref(C.c_void_p())) == -1
    assert lib.dll.rb200_gather_plan(None, 0, 0, None) == -1
