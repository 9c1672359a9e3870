"""C-ABI: the shared library loads and exports exactly what include/sonet_b200.h declares;
argument validation and the host (CPU) plugin entry points work without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sonet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sonet_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib_built):
    h = ctypes.CDLL(lib_built)
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(h, n), "library does not export %s" % n


def test_python_binding_covers_header(lib_built):
    from sonet_b200 import _C
    assert sorted(_C.EXPORTED_SYMBOLS) == declared_symbols()
    lib = _C.lib()
    assert b"sm_100a" in lib.sonet_version()


def test_library_has_no_torch_dependency(lib_built):
    import subprocess
    out = subprocess.run(["ldd", lib_built], capture_output=True, text=True).stdout
    assert "libtorch" not in out and "libc10" not in out


def test_bad_arguments_return_error_codes(lib_built):
    from sonet_b200 import _C
    lib = _C.lib()
    # K out of range is rejected before any CUDA call
    rc = lib.sonet_index_max_f32(None, None, 1, 1, 1, 0, None, None, None)
    assert rc == -1 and "K=0" in _C.last_error()
    rc = lib.sonet_index_max_f32(None, None, 1, 1, 4, 64, None, None, None)
    assert rc == -1 and "null" in _C.last_error()
    rc = lib.sonet_som_assign(None, None, 1, 8, 500, 3, None, None, None, None, None, None)
    assert rc == -1 and "M=500" in _C.last_error()
    rc = lib.sonet_som_assign(None, None, 1, 8, 64, 7, None, None, None, None, None, None)
    assert rc == -1
    rc = lib.sonet_chamfer_f32(None, None, 0, 1, 1, None, None, None, None, None, None, None, None)
    assert rc == -1
    with pytest.raises(RuntimeError):
        _C.check(rc, "chamfer")
    # empty batches are a no-op success
    assert lib.sonet_index_max_f32(None, None, 0, 4, 4, 8, None, None, None) == 0
    assert lib.sonet_pointwise_layer_f32(None, 4, None, 0, 0, 16, None, None, None, 8, 1, None,
                                         None, 0, None, None) == 0


def test_cuda_entry_points_reject_host_tensors():
    from sonet_b200 import index_max
    data = torch.zeros(1, 2, 8)
    idx = torch.zeros(1, 8, dtype=torch.int32)
    with pytest.raises(RuntimeError, match="CUDA"):
        index_max.forward_cuda(data, idx, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        index_max.forward_cuda_shared_mem(data, idx, 4)


def test_plugin_cpu_entry_points_known_answer(oracle_mod):
    """SURVEY.md §8c known-answer fact, verified on the reference binary."""
    from sonet_b200 import index_max
    d = torch.tensor([[[1, 5, 5, -2000, 3, 3]]], dtype=torch.float32)
    i = torch.tensor([[0, 1, 1, 2, 3, 3]], dtype=torch.int32)
    assert index_max.forward_cpu(d, i, 5).tolist() == [[[0, 1, 0, 4, 0]]]
    assert index_max.forward_multi_thread_cpu(d, i, 5, 3).tolist() == [[[0, 1, 0, 4, 0]]]
    rs = np.random.RandomState(0)
    data = torch.from_numpy(rs.normal(size=(3, 21, 257)).astype(np.float32))
    index = torch.from_numpy(rs.randint(0, 13, size=(3, 257)).astype(np.int32))
    want = oracle_mod.index_max(data, index, 13)
    assert torch.equal(index_max.forward_cpu(data, index, 13), want)
    assert torch.equal(index_max.forward_multi_thread_cpu(data, index, 13, 4), want)
    with pytest.raises(RuntimeError):
        index_max.forward_cpu(data, index.clamp(min=0) + 100, 13)
