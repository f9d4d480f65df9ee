"""CPU: the C-ABI library builds/loads, exports every symbol include/sdvg.h declares, and fails LOUDLY
without a GPU (no CPU fallback anywhere in the product path)."""
import ctypes as C
import os
import re

import pytest
import torch

from conftest import ROOT


def header_functions():
    src = open(os.path.join(ROOT, "include", "sdvg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sdvg_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from sdvg_b200 import _lib
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"libsdvg.so does not export {n}"
    assert sorted(_lib.SYMBOLS) == names, "ctypes binding list and header disagree"
    assert lib.sdvg_version() == 100


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "sd-video-gen_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                # no import, path or attribute use of the oracle package (prose mentions in comments are fine)
                assert not re.search(r"(from|import)\s+oracle\b|\boracle[./]\w|\boracle\s*=|#include.*oracle", text), \
                    f"{f} uses oracle/"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    from sdvg_b200 import _lib
    lib = _lib.load()
    cfg = _lib.SdvgConfig(32, 4, 1, 1, 256, 2048, 1e-5, 4, 8, 16, 1, 0)
    h = C.c_void_p()
    rc = lib.sdvg_create(C.byref(cfg), C.byref(h))
    assert rc == -2 and not h.value                       # SDVG_ERR_CUDA
    assert b"no CPU fallback" in lib.sdvg_last_error(None)


def test_module_raises_on_cpu_tensors():
    import sdvg_b200
    m = sdvg_b200.Transformer(0, 32, 4, 1, 1, 0.1, frame_size=64).eval()
    x = torch.zeros(2, 3, 256)
    with pytest.raises(RuntimeError, match="CUDA only"):
        m(x, x)
    with pytest.raises(RuntimeError, match="CUDA only"):
        sdvg_b200.rollout(m, x, 2)
    with pytest.raises(RuntimeError):
        sdvg_b200.gemm(torch.zeros(4, 8), torch.zeros(4, 8))


def test_bad_config_is_rejected_before_touching_cuda():
    from sdvg_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    cfg = _lib.SdvgConfig(30, 4, 1, 1, 256, 2048, 1e-5, 4, 8, 16, 1, 0)   # d % H != 0
    assert lib.sdvg_create(C.byref(cfg), C.byref(h)) == -1
    cfg = _lib.SdvgConfig(32, 4, 1, 1, 256, 2048, 1e-5, 4, 64, 16, 1, 0)  # max_tokens > 32
    assert lib.sdvg_create(C.byref(cfg), C.byref(h)) == -1
