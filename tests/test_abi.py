"""CPU: the C-ABI library loads, exports every symbol include/gnnx.h declares, and the product
path fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest
import torch

import gnnx
from gnnx import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "gnnx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert len(syms) >= 15
    lib = ctypes.CDLL(_abi.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), "libgnnx.so does not export %s" % s
    assert set(_abi.EXPORTS) == set(syms), "python binding and header disagree"
    assert _abi.lib().gx_version() == _abi.GX_VERSION == 211


def test_struct_layout_matches_defaults():
    hp = _abi.GxHparams()
    _abi.lib().gx_default_hparams(ctypes.byref(hp))
    assert hp.num_epochs == 100 and abs(hp.lr - 0.1) < 1e-8 and abs(hp.beta2 - 0.999) < 1e-7
    assert abs(hp.coef_size - 0.005) < 1e-9 and hp.coef_lap == 1.0 and hp.init == _abi.GX_INIT_M0
    assert hp.seed == 0 and hp.start_step == 0 and hp.opt == 0 and hp.opt_scheduler == 0 and hp.opt_decay_rate == 1.0 and ctypes.sizeof(hp) == 80
    # gx_explain_io: twelve pointers, in the header's order
    src = open(os.path.join(ROOT, "include", "gnnx.h")).read()
    body = src[src.index("typedef struct gx_explain_io {"):src.index("} gx_explain_io;")]
    names = re.findall(r"\b(?:const\s+)?float\*\s+([a-z0-9_]+);", body)
    assert names == [n for n, _ in _abi.GxExplainIo._fields_]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    with pytest.raises(_abi.GnnxError) as ei:
        gnnx.Engine(0)
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "gnn-model-explainer_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "gnnx_oracle" not in txt and "ref_harness" not in txt, "%s touches the oracle" % f
