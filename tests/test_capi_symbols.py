"""The C-ABI library loads on a CPU-only box and exports every symbol include/espnet_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "espnet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(espb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_bound_symbols():
    from espnet_b200 import lib

    assert _header_symbols() == sorted(lib.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    from espnet_b200 import lib

    if not os.path.exists(lib.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    handle = ctypes.CDLL(lib.LIB_PATH)
    for sym in _header_symbols():
        assert hasattr(handle, sym), sym
    handle.espb_abi_version.restype = ctypes.c_int
    assert handle.espb_abi_version() == 1


def test_ops_refuse_cpu_tensors():
    """No CPU fallback: the product path fails loudly instead of computing on the host."""
    import torch

    import espnet_b200

    fe = espnet_b200.DefaultFrontend()
    with pytest.raises((AssertionError, RuntimeError)):
        fe(torch.zeros(1, 4000), torch.tensor([4000]))
    with pytest.raises(RuntimeError):
        espnet_b200.Speech2Text(asr_model=None, device="cpu")


def test_state_dict_names_match_reference_fixture():
    import argparse
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import refbuild
    from golden_util import load

    import espnet_b200

    for case in ("tiny", "small"):
        z, cfg, w = load(case)
        model = espnet_b200.build_model(argparse.Namespace(**refbuild.model_yaml(cfg)))
        model.load_state_dict(w, strict=True)


def test_unsupported_configs_are_rejected():
    import espnet_b200

    with pytest.raises(NotImplementedError):
        espnet_b200.ConformerEncoder(80, 256, rel_pos_type="legacy", macaron_style=True)
    with pytest.raises(NotImplementedError):
        espnet_b200.DefaultFrontend(n_fft=400)
