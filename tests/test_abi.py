"""CPU: the C-ABI library builds, loads and exports every symbol include/mp3gpu.h declares; argument
validation that needs no GPU behaves; and without a GPU the product path FAILS LOUDLY (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(pkg):
    if not os.path.exists(pkg.host.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return pkg.load_library()


def declared_functions():
    src = open(os.path.join(ROOT, "include", "mp3gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mp3gpu_[A-Za-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib, pkg):
    names = declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(pkg.host.EXPORTS) == names


def test_legacy_entry_points_exported(lib, pkg):
    """include/mp3gpu_legacy.h: the reference's own five symbols (musicin.c:754-779) + shim control"""
    src = open(os.path.join(ROOT, "include", "mp3gpu_legacy.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = re.findall(r"^(?:void|int|long)\s+([A-Za-z0-9_]+)\s*\(", src, flags=re.M)
    assert sorted(protos) == sorted(pkg.host.LEGACY_EXPORTS)
    for n in protos:
        assert hasattr(lib, n), n


def test_struct_layouts(pkg):
    assert pkg.host.PSY_DT.itemsize == 8 + 21 * 8 + 36 * 8 + 8
    assert pkg.host.FO_DT.itemsize == 16


def test_argument_validation_and_no_cpu_fallback(lib, pkg):
    import torch
    cfg = pkg.host.Config(22050, 2, 128, 1, 1, 0)           # LSF rate: the reference refuses it too (l3psy.c:169-176)
    ctx = C.c_void_p()
    assert lib.mp3gpu_create(C.byref(cfg), C.byref(ctx)) == -1
    assert b"sampling" in lib.mp3gpu_last_error()
    cfg = pkg.host.Config(44100, 3, 128, 1, 1, 0)
    assert lib.mp3gpu_create(C.byref(cfg), C.byref(ctx)) == -1
    cfg = pkg.host.Config(44100, 2, 100, 1, 1, 0)
    assert lib.mp3gpu_create(C.byref(cfg), C.byref(ctx)) == -1
    if not torch.cuda.is_available():
        with pytest.raises(pkg.Mp3GpuError, match="no CPU fallback|CUDA"):
            pkg.Encoder(44100, 2, 128)
