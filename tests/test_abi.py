"""The C-ABI library exists, loads, and exports every symbol include/gnas.h declares (no compute)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "gnas.h")).read()
    return sorted(set(re.findall(r"\b(gnas_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_bound_by_the_python_side():
    from genome_nas_b200 import _lib
    assert sorted(_lib.EXPORTS) == _declared()


def test_cuda_library_exports_every_declared_symbol():
    import __graft_entry__ as entry
    entry.build()
    from genome_nas_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH)
    try:
        lib = ctypes.CDLL(_lib.LIB_PATH)
    except OSError as e:  # libcudart missing on a box without the toolkit is an environment error
        raise AssertionError(f"libgnas.so does not load: {e}")
    for name in _declared():
        assert hasattr(lib, name), name
    lib.gnas_version.restype = ctypes.c_int
    assert lib.gnas_version() >= 100


def test_descriptor_layout_matches_header():
    from genome_nas_b200 import _lib
    # 14 int32/float scalars, 14*9 switches, padding to 8, then int64 tables
    assert ctypes.sizeof(_lib.CellDesc) == 56 + 128 + 8 * (14 * 9 * 4 + 2 + 14 * 9 * 2 + 2)
    assert ctypes.sizeof(_lib.RhnDesc) == 24 + 180
