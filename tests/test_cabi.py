"""The C-ABI library loads and exports every symbol include/vlfs_b200.h declares; without a
GPU the context constructor fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import pytest
import vlfs_b200
from vlfs_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "vlfs_b200.h")).read()
    return sorted(set(re.findall(r"\b(vlfs_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared() == sorted(L.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = L.load()
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.vlfs_version() >= 100


def test_struct_layouts_match_header_sizes():
    # 4 int32 + 9 pointers ; see vlfs_slot_desc
    assert C.sizeof(L.SlotDesc) == 16 + 9 * 8
    assert C.sizeof(L.TermDesc) == 8 + 16 + 8 + 16 + 16 + 24 + 24
    assert C.sizeof(L.SolveStats) == 8 + 8 + 16


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(L.VlfsError):
        vlfs_b200.B200System(10, False)
