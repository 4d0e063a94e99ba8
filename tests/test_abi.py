"""CPU: the C-ABI library builds, loads, and exports every symbol include/usflow_b200.h declares
(no compute calls -- there is no GPU here), and the product refuses CPU tensors (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "usflow_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(usf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from nf4ad_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build with python -m nf4ad_b200.build"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, missing
    # every declared function is also bound with a prototype on the Python side
    unbound = [n for n in names if n not in _lib.EXPORTED]
    assert not unbound, unbound
    assert _lib.lib().usf_version() == 100


def test_no_torch_types_in_abi():
    src = open(os.path.join(ROOT, "include", "usflow_b200.h")).read()
    assert "torch" not in src.replace("no torch types", "").replace("There are no\n * torch types", "")
    assert "at::Tensor" not in src


def test_cpu_tensors_are_refused():
    import nf4ad_b200
    from _cases import build_flow
    flow = build_flow(nf4ad_b200.namespace(), "NonUSFlow", 8, 2, ("mlp", [16]), affine_conjugation=True)
    with pytest.raises(nf4ad_b200.USFError):
        flow.log_prob(torch.randn(3, 8))
    with pytest.raises(nf4ad_b200.USFError):
        flow.backward(torch.randn(3, 8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "nf4ad_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, os.path.join(dirpath, f)


def test_state_dict_keys_match_oracle(O):
    import nf4ad_b200
    from _cases import build_flow
    ns = nf4ad_b200.namespace()
    for kind, cond, kw in (("NonUSFlow", ("mlp", [16]), dict(affine_conjugation=True, prior_scale=1.0)),
                           ("USFlow", ("densenn1", [16, 8]), dict(affine_conjugation=False, householder=0)),
                           ("NonUSFlow", ("densenn2", [8]), dict(lu_transform=2, householder=2))):
        a = build_flow(O, kind, 6, 3, cond, **kw)
        b = build_flow(ns, kind, 6, 3, cond, **kw)
        assert list(a.state_dict().keys()) == list(b.state_dict().keys())
        b.load_state_dict(a.state_dict())
