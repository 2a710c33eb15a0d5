"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol the
header declares, and rejects bad arguments with error codes (no compute call needs a GPU here)."""
import ctypes
import os

import pytest

from zhilight_b200 import _lib


def test_library_exports_every_declared_symbol(lib):
    protos = _lib.parse_header()
    assert len(protos) >= 40
    for name in protos:
        assert hasattr(lib, name), name


def test_header_cites_reference_for_compute_entry_points():
    src = open(_lib.HEADER).read()
    for must in ("q_gemm_k_major.cu", "attention_kernel.cu", "rotary_embedding_fuse_cache.cu", "layernorm.cu",
                 "ragged_buffer_kernel.cu", "linear.cpp", "utils.cu"):
        assert must in src, must


def test_version_and_error_string(lib):
    assert lib.zl_version() >= 100
    rc = lib.zl_gptq_increase_zero(None, 0, None)
    assert rc == -1
    assert b"invalid argument" in lib.zl_last_error()


def test_packed_bytes(lib):
    assert lib.zl_w4_packed_bytes(4096, 4096, 128) == (4096 // 32) * 32 * 2128
    assert lib.zl_w4_packed_bytes(4096, 4096, 64) == 0          # unsupported group size
    assert lib.zl_w4_packed_bytes(100, 4096, 128) == 0          # N % 32


def test_argument_validation_without_gpu(lib):
    p = ctypes.c_void_p(16)
    # group size != 128 is reported as unsupported before any CUDA call
    assert lib.zl_w4a16_gemm(p, 4096, p, None, None, p, 1, 4096, 4096, 64, 0, 0, None) == -2
    # misaligned leading dimension
    assert lib.zl_w4a16_gemm(p, 4097, p, None, None, p, 1, 4096, 4096, 128, 0, 0, None) == -1
    # residual epilogue without residual pointer
    assert lib.zl_w4a16_gemm(p, 4096, p, None, None, p, 1, 4096, 4096, 128, 2, 0, None) == -1
    assert lib.zl_decode_attention(p, p, p, p, None, 1.0, 128, p, 1, 1, 32, 8, 96, 1, None, 0, 0, 0, None) == -2
    assert lib.zl_rmsnorm(p, p, p, 1, 4095, 1e-5, 1.0, 0, 0, None) == -2


def test_ops_refuse_cpu_tensors():
    import torch
    from zhilight_b200 import ops
    with pytest.raises(_lib.ZLError):
        ops.rmsnorm(torch.zeros(1, 64, dtype=torch.float16), torch.ones(64, dtype=torch.float16), 1e-5)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.ZLError):
        _lib.load()
