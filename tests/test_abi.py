"""CPU: the C-ABI library loads, exports every symbol the header declares, the ctypes table matches
the header, and the product path refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from jyutvoice_b200 import _lib, build as jbuild


@pytest.fixture(scope="module")
def built():
    jbuild.build()  # no-op when the .so is fresh
    return _lib.lib()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "jyutvoice_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jv_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built):
    syms = header_symbols()
    assert len(syms) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in the header but not exported"


def test_ctypes_table_matches_header():
    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_version_and_error_string(built):
    assert built.jv_version() >= 1
    assert isinstance(built.jv_last_error(), bytes)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback(built):
    h = ctypes.c_void_p()
    rc = built.jv_estimator_create(0, 1, ctypes.byref(h))
    assert rc != 0 and h.value is None
    with pytest.raises(RuntimeError):
        _lib.check(rc)
    from jyutvoice_b200 import CausalConditionalCFM, HiFTGenerator
    cfm = CausalConditionalCFM()
    with pytest.raises(RuntimeError):
        cfm(torch.zeros(1, 80, 8), torch.ones(1, 1, 8), 2, spks=torch.zeros(1, 80))
    with pytest.raises(RuntimeError):
        HiFTGenerator().inference(torch.zeros(1, 80, 8))


def test_state_dict_keys_match_oracle_tables():
    from jyutvoice_b200 import CausalConditionalCFM, HiFTGenerator
    from jyutvoice_b200 import synthetic
    cfm = CausalConditionalCFM()
    want = {k: tuple(s) for k, s, _ in synthetic.estimator_table()}
    got = {k: tuple(v.shape) for k, v in cfm.state_dict().items()}
    assert got == want
    hift = HiFTGenerator()
    got = {k: tuple(v.shape) for k, v in hift.state_dict().items()}
    want = {k: tuple(v.shape) for k, v in synthetic.make_hift_state_dict().items()}
    assert got == want


def test_reference_config_guard():
    from jyutvoice_b200 import CausalConditionalDecoder, HiFTGenerator
    with pytest.raises(ValueError):
        CausalConditionalDecoder(channels=(256, 256))
    with pytest.raises(ValueError):
        HiFTGenerator(upsample_rates=(8, 8))


def test_mask_to_lengths():
    from jyutvoice_b200.flow_matching import _lens_from_mask
    m = torch.zeros(2, 1, 5)
    m[0, 0, :5] = 1
    m[1, 0, :2] = 1
    assert _lens_from_mask(m) == [5, 2]
    m[1, 0, 4] = 1  # hole: not a prefix mask
    with pytest.raises(ValueError):
        _lens_from_mask(m)
