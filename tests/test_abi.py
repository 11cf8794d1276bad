"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/jsfe.h declares,
and refuses to work (loudly) without a CUDA device -- there is no CPU fallback behind it."""
import ctypes as C
import os
import re

import pytest
import torch

from conftest import ROOT
from jetson_slam_b200 import frontend


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "jsfe.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(jsfe_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(frontend.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/jsfe.h but not exported by libjsfe.so"


def test_product_does_not_reference_the_oracle():
    """The product tree must not import, include or link anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "jetson_slam_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.replace("the oracle", "").replace("(oracle/)", "").replace("oracle (", "") or f == "__init__.py" or \
                    not re.search(r"(import|include|from)\s+[\"<]?\.*oracle", src), f
    for d in ("compat",):
        p = os.path.join(ROOT, d)
        if os.path.isdir(p):
            for dirpath, _, files in os.walk(p):
                for f in files:
                    assert not re.search(r"(import|include|from)\s+[\"<]?\.*oracle", open(os.path.join(dirpath, f), errors="ignore").read()), f


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a GPU-less host")
def test_create_fails_loudly_without_gpu():
    with pytest.raises(frontend.JsfeError) as e:
        frontend.Frontend(120, 160, n_levels=3, tile_h=16, tile_w=16)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_bad_arguments_are_rejected_before_touching_cuda():
    lib = frontend.lib()
    assert lib.jsfe_create(None, None) == -1
    assert b"null" in lib.jsfe_last_error()
    assert lib.jsfe_max_keypoints(None) == -1
