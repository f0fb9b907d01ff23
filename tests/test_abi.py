"""CPU-side checks of the C-ABI boundary: the shared library loads, exports every symbol that
include/yv6.h declares, and fails loudly (no fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from yolov6_b200 import _lib


def declared_symbols():
    with open(os.path.join(ROOT, "include", "yv6.h")) as f:
        src = f.read()
    return sorted(set(re.findall(r"\b(yv6_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/yv6.h but not exported"
    assert set(_lib.exported_symbols()) <= set(names)
    assert lib.yv6_abi_version() == 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    with pytest.raises(RuntimeError):
        _lib.handle(0)
    from yolov6_b200.model import build_model
    from yolov6_b200.nms import non_max_suppression
    m = build_model("yolov6n", 80, torch.device("cpu")).eval()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError):
        non_max_suppression(torch.zeros(1, 10, 85))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "yolov6_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{fn} imports the oracle"


def test_ctypes_mirrors_match_the_header_struct_sizes(tmp_path):
    """The ctypes Structure mirrors in yolov6_b200/_lib.py must have the sizes a C compiler gives the structs
    of include/yv6.h (caught a silent field drift once)."""
    import ctypes as C
    import shutil
    import subprocess
    from yolov6_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "yv6.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(yv6_conv_desc), '
                   'sizeof(yv6_stem_desc), sizeof(yv6_loss_desc), sizeof(yv6_wgrad_desc), sizeof(yv6_bn_desc), '
                   'sizeof(yv6_bn_stats_desc), sizeof(yv6_xform_seg));return 0;}\n')
    exe = tmp_path / "sz"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.run(["gcc", "-I", inc, str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [C.sizeof(_lib.ConvDesc), C.sizeof(_lib.StemDesc), C.sizeof(_lib.LossDesc), C.sizeof(_lib.WgradDesc), C.sizeof(_lib.BnDesc),
            C.sizeof(_lib.BnStatsDesc), C.sizeof(_lib.XformSeg)]
    assert got == want, (got, want)
