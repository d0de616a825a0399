"""The C-ABI library loads and exports every function include/bevbert_b200.h declares (no compute calls)."""
import os
import re

from bevbert_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "bevbert_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bb_[a-z0-9_]+)\s*\(", src)))


def test_header_binding_and_library_agree():
    names = _declared()
    assert len(names) >= 25
    assert sorted(_lib.EXPORTED_SYMBOLS) == names, set(names) ^ set(_lib.EXPORTED_SYMBOLS)
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), n
    assert lib.bb_abi_version() == 1
    assert lib.bb_launch_count() >= 0


def test_gemm_args_struct_layout():
    import ctypes
    assert ctypes.sizeof(_lib.GemmArgs) == 216
    # values printed by a C program using sizeof / offsetof on include/bevbert_b200.h
    assert ctypes.sizeof(_lib.AttnDesc) == 280 and _lib.AttnDesc.seed_attn.offset == 128 and _lib.AttnDesc.dbias.offset == 272
    assert ctypes.sizeof(_lib.PanoDesc) == 320 and _lib.PanoDesc.seed_attn.offset == 136 and _lib.PanoDesc.dbe2.offset == 312
    assert ctypes.sizeof(_lib.FfnDesc) == 184 and _lib.FfnDesc.dbeta.offset == 176
    F = _lib.FlashArgs
    assert ctypes.sizeof(F) == 248 and (F.lse.offset, F.dout.offset, F.dsum.offset, F.dq.offset, F.lddq.offset,
                                        F.dbias.offset) == (120, 144, 168, 176, 224, 240)
    assert _lib.GemmArgs.block_n.offset == 208 and _lib.GemmArgs.add_in.offset == 200  # == sizeof/offsetof in C
