"""CPU tests of the boundary: the C-ABI library loads and exports what include/sg_b200.h declares."""
import ctypes as C
import os
import re

import pytest

from scenario_gym_b200 import abi

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(REPO, "include", "sg_b200.h")).read()
    return sorted(set(re.findall(r"\b(sg_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    assert {"sg_reset", "sg_rollout", "sg_rollout_host", "sg_test_box_pairs", "sg_sizeof",
            "sg_abi_version", "sg_last_error", "sg_default_params"} <= set(syms)


def test_product_library_exports_every_declared_symbol():
    if not os.path.exists(abi.PRODUCT_LIB):
        import __graft_entry__ as g

        g.build()
    lib = C.CDLL(abi.PRODUCT_LIB)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/sg_b200.h but not exported"
    f = abi.bind(lib, "sg_")  # checks ABI version + struct sizes; no compute call
    p = abi.SgParams()
    f["default_params"](C.byref(p))
    d = abi.default_params()
    for name, _ in abi.SgParams._fields_:
        a, b = getattr(p, name), getattr(d, name)
        assert a == b or (a != a and b != b), name


def test_oracle_exports_same_abi(oracle_lib):
    p = abi.SgParams()
    oracle_lib["default_params"](C.byref(p))
    d = abi.default_params()
    for name, _ in abi.SgParams._fields_:
        a, b = getattr(p, name), getattr(d, name)
        assert a == b or (a != a and b != b), name


def test_engine_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from scenario_gym_b200 import synthetic
    from scenario_gym_b200.engine import Engine

    scene = synthetic.pack_synthetic(synthetic.vehicles_config(0, N=2, M=4, T=4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(scene)
