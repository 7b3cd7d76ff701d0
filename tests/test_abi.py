"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/maestro_b200.h declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from maestro_b200 import abi, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "maestro_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mgpu_[a-zA-Z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    l = lib.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(l, s), "libmaestro_b200.so does not export %s" % s


def test_python_abi_mirror_matches_header():
    assert sorted(abi.all_symbols()) == header_symbols()


def test_struct_layout_matches_c():
    # sizes computed by hand from the header: fab = 8 + 12 + 12 + 4 + 4 + 12 = 52 -> padded to 56
    assert C.sizeof(abi.mgpu_fab) == 56
    assert C.sizeof(abi.mgpu_params) == 4 * 28 + 8 * 6


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    l = lib.load()
    assert l.mgpu_init(0) != 0
    assert b"no CUDA device" in l.mgpu_last_error()
    from maestro_b200 import Fab, face_fabs, make_params, make_adv_bc

    p = make_params(2, n=[8, 8, 1])
    s = Fab([0, 0, 0], [7, 7, 0], 4, p.nscal, dm=2)
    f = Fab([0, 0, 0], [7, 7, 0], 1, p.nscal, dm=2)
    with pytest.raises(RuntimeError, match="not initialised"):
        lib.ops().make_edge_scal(p, s, face_fabs([0, 0, 0], [7, 7, 0], 0, p.nscal, 2),
                                 face_fabs([0, 0, 0], [7, 7, 0], 1, 1, 2), f,
                                 make_adv_bc(p, [[-1, -1], [-1, -1]]), False, 1, 3, 1, False)
