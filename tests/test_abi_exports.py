"""The C-ABI library loads here (no GPU, no compute calls) and exports every symbol include/ps_b200.h declares."""
import ctypes
import os
import re

from common import ROOT


def _declared():
    txt = open(os.path.join(ROOT, "include", "ps_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ps_[a-z_0-9]+)\s*\(", txt)))


def test_header_declares_the_boundary():
    names = _declared()
    for must in ("ps_init", "ps_batch_create", "ps_batch_destroy", "ps_batch_step", "ps_batch_get_state", "ps_batch_set_state",
                 "ps_batch_get_contacts", "ps_batch_get_candidates", "ps_batch_apply_impulse", "ps_batch_add_body", "ps_batch_reset",
                 "ps_batch_stats", "ps_last_error", "ps_broadphase_pairs"):
        assert must in names


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(g.LIB)
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing
    from physicssimulator_b200 import _native
    assert sorted(_native.EXPORTS) == _declared()


def test_no_signature_mentions_torch():
    txt = open(os.path.join(ROOT, "include", "ps_b200.h")).read()
    assert "torch" not in txt.lower() and "at::" not in txt


def test_config_struct_layout_matches_ctypes():
    from physicssimulator_b200.configuration import ps_step_config
    assert ctypes.sizeof(ps_step_config) == 5 * 8 + 10 * 4 + 6 * 8
    from physicssimulator_b200._native import ps_stats
    assert ctypes.sizeof(ps_stats) == 14 * 8


def test_product_never_touches_the_oracle():
    """a product path that routes through oracle/ (or any CPU fallback) voids every parity claim"""
    pkg = os.path.join(ROOT, "physicssimulator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "libps_oracle" not in src, f
                assert "oracle/" not in src.replace("the oracle", ""), f


def test_missing_gpu_is_loud_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        return
    from physicssimulator_b200 import _native
    L = _native.lib()
    rc = L.ps_init(ctypes.c_int32(-1))
    assert rc == _native.PS_ERR_CUDA
    assert b"no CUDA device" in L.ps_last_error()
