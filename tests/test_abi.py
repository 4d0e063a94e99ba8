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


def test_host_narrowing_is_round_to_nearest_even():
    """usf_host_f32_to_bf16 (host cores, no CUDA): the bit pattern of a round-to-nearest-even fp32 -> bf16 conversion, for
    contiguous and strided rows, any thread count, specials included -- what the device's first step would produce, so that
    narrowing before the PCIe copy cannot change a score."""
    import ctypes as C
    import torch
    from nf4ad_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1237, 83, generator=g) * torch.logspace(-30, 30, 83)
    x[0, :8] = torch.tensor([float("inf"), -float("inf"), 0.0, -0.0, 1e-40, -1e-45, 3.3895e38, 3.4e38])
    x[1, 0] = float("nan")
    # ties: exactly half way between two bf16 values, both parities
    x[2, 0] = torch.tensor([0x3F808000], dtype=torch.int32).view(torch.float32)[0]
    x[2, 1] = torch.tensor([0x3F818000], dtype=torch.int32).view(torch.float32)[0]
    ref = x.to(torch.bfloat16)
    for threads in (1, 3, 8, 0):
        for ldd in (83, 96):
            out = torch.full((1237, ldd), 7.0, dtype=torch.bfloat16)
            assert L.usf_host_f32_to_bf16(C.c_void_p(x.data_ptr()), 83, C.c_void_p(out.data_ptr()), ldd, 1237, 83, threads) == 0
            got = out[:, :83]
            nan = torch.isnan(ref)
            assert torch.equal(torch.isnan(got), nan)
            assert torch.equal(got.view(torch.int16)[~nan], ref.view(torch.int16)[~nan])
            assert bool((out[:, 83:] == 7.0).all())            # pad columns untouched
    # strided source rows, empty inputs, bad arguments
    xs = x[:, :40]
    out = torch.empty(1237, 40, dtype=torch.bfloat16)
    assert L.usf_host_f32_to_bf16(C.c_void_p(xs.data_ptr()), 83, C.c_void_p(out.data_ptr()), 40, 1237, 40, 2) == 0
    nan = torch.isnan(ref[:, :40])
    assert torch.equal(out.view(torch.int16)[~nan], ref[:, :40].contiguous().view(torch.int16)[~nan])
    assert L.usf_host_f32_to_bf16(None, 83, None, 83, 0, 83, 1) == 0
    # the plain staged copy of the same pool (pageable fp32 rows -> pinned ring)
    cp = torch.full((1237, 48), 7.0)
    assert L.usf_host_copy_f32(C.c_void_p(xs.data_ptr()), 83, C.c_void_p(cp.data_ptr()), 48, 1237, 40, 3) == 0
    assert torch.equal(cp[:, :40].view(torch.int32), xs.contiguous().view(torch.int32)) and bool((cp[:, 40:] == 7.0).all())
    # a forked child (e.g. a data-loader worker) has none of the parent's pool threads: it must get its own pool.
    # No torch calls in the child (its OpenMP pool does not survive a fork either); the parent waits with a deadline.
    import os
    import signal
    import time
    import warnings
    x2 = torch.randn(500, 64, generator=g)
    want = bytes(x2.to(torch.bfloat16).view(torch.int16).numpy().tobytes())
    src_ptr = C.c_void_p(x2.data_ptr())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)
        pid = os.fork()
    if pid == 0:
        try:
            buf = (C.c_uint16 * (500 * 64))()
            rc = L.usf_host_f32_to_bf16(src_ptr, 64, buf, 64, 500, 64, 4)
            os._exit(0 if rc == 0 and bytes(buf) == want else 1)
        finally:
            os._exit(2)
    deadline = time.time() + 30
    status = None
    while time.time() < deadline:
        done, st = os.waitpid(pid, os.WNOHANG)
        if done == pid:
            status = st
            break
        time.sleep(0.05)
    if status is None:
        os.kill(pid, signal.SIGKILL)
        os.waitpid(pid, 0)
        raise AssertionError("forked child hung in usf_host_f32_to_bf16")
    assert os.WIFEXITED(status) and os.WEXITSTATUS(status) == 0
    assert L.usf_host_f32_to_bf16(C.c_void_p(x.data_ptr()), 10, C.c_void_p(out.data_ptr()), 40, 5, 40, 1) != 0


def test_lu_log_prior_matches_the_oracle_formula():
    """`LUTransform.log_prior` is written with fixed-shape ops (graph-capturable); same value as the oracle's
    boolean-indexed form (oracle/shim/src/usflows/transforms.py)."""
    import torch
    import oracle
    from nf4ad_b200.transforms import LUTransform
    ns = oracle.load()
    torch.manual_seed(3)
    ours = LUTransform(9, prior_scale=0.7).double()
    theirs = ns.transforms.LUTransform(9, prior_scale=0.7).double()
    theirs.load_state_dict(ours.state_dict())
    a, b = ours.log_prior(), theirs.log_prior()
    assert abs(float(a) - float(b)) <= 1e-12 * abs(float(b))
    assert LUTransform(4, prior_scale=None).log_prior() == 0.0


def test_fused_adam_has_no_cpu_path():
    """`FusedAdam` validates its hyper-parameters like torch's Adam and refuses CPU parameters (no fallback)."""
    import pytest
    import torch
    from nf4ad_b200 import _lib
    from nf4ad_b200.optim import FusedAdam
    p = torch.nn.Parameter(torch.zeros(4))
    for bad in (dict(lr=-1.0), dict(betas=(1.0, 0.9)), dict(eps=-1e-8), dict(weight_decay=-0.1)):
        with pytest.raises(ValueError):
            FusedAdam([p], **bad)
    opt = FusedAdam([p], lr=1e-3)
    assert all(g["capturable"] for g in opt.param_groups)        # what DataParallelTrainer checks before capturing a step
    p.grad = torch.ones(4)
    with pytest.raises(_lib.USFError):
        opt.step()


def _header_prototypes():
    """name -> (return class, [argument classes]) parsed from include/usflow_b200.h.  Classes: 'p' pointer (incl. the
    stream handle), 'i64', 'i32', 'f32', 'f64', 'sz', 'str' (const char* return)."""
    src = open(os.path.join(ROOT, "include", "usflow_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)

    def cls(decl, is_return=False):
        d = " ".join(decl.replace("const", " ").split())
        if "*" in d:
            return "str" if (is_return and d.startswith("char")) else "p"
        base = d.split(" ")[0] if not d.startswith("unsigned") else " ".join(d.split(" ")[:2])
        table = {"int64_t": "i64", "int": "i32", "int32_t": "i32", "float": "f32", "double": "f64", "size_t": "sz",
                 "usf_stream_t": "p", "long": "i64"}
        assert base in table, f"unclassified C type in the header: {decl!r}"
        return table[base]

    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(usf_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.startswith("typedef") or not ret:
            continue
        arglist = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        protos[name] = (cls(ret, True), [cls(a) for a in arglist])
    return protos


def test_ctypes_prototypes_match_the_header():
    """Every entry point's ctypes prototype in `nf4ad_b200/_lib.py` agrees with its C declaration argument by argument
    (count and class: pointer / int64 / int / float / size_t) -- a float passed where the C side reads an int64, or a
    missing argument, would not fail loudly at the call."""
    from nf4ad_b200 import _lib
    hdr = _header_prototypes()
    assert set(hdr) == set(_declared())

    def cls(t):
        if t is None:
            return "void"
        if t is ctypes.c_char_p:
            return "str"
        if t is ctypes.c_void_p or (isinstance(t, type) and issubclass(t, ctypes._Pointer)):
            return "p"
        return {ctypes.c_int64: "i64", ctypes.c_longlong: "i64", ctypes.c_int32: "i32", ctypes.c_int: "i32",
                ctypes.c_float: "f32", ctypes.c_double: "f64", ctypes.c_size_t: "sz"}[t]

    bad = {}
    for name, (ret, args) in hdr.items():
        pres, pargs = _lib._PROTOS[name]
        got = (cls(pres), [cls(a) for a in pargs])
        want = (ret, args)
        if got[0] == "str" and want[0] in ("str", "p"):
            got = (want[0], got[1])
        # a `const char*` ARGUMENT (c_char_p) is a pointer
        got = (got[0], ["p" if a == "str" else a for a in got[1]])
        if got != want:
            bad[name] = {"header": want, "ctypes": got}
    assert not bad, bad


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """The descriptor structs `_lib.py` / `optim.py` hand to the library have the C compiler's own layout: a tiny C program
    including include/usflow_b200.h (plain C: the header must stay C-compatible) prints sizeof / offsetof of every field."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc") or shutil.which("cc")
    if gcc is None:
        pytest.skip("no C compiler")
    from nf4ad_b200 import _lib
    from nf4ad_b200.optim import _AdamTensor
    structs = {"usf_linear_desc": (_lib.LinearDesc, ["W", "Wb", "bias", "N", "K", "ldw"]),
               "usf_block_desc": (_lib.BlockDesc, ["G", "b_off", "n_mlp", "mlp", "Da", "Db", "C", "affine", "clamp"]),
               "usf_stack_desc": (_lib.StackDesc, ["D", "n_blocks", "blocks", "G_final", "inverse", "base_kind", "loc",
                                                   "inv_scale", "const_term", "ctx_dim"]),
               "usf_adam_tensor": (_AdamTensor, ["p", "g", "m", "v", "n"])}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "usflow_b200.h"', "int main(void) {"]
    for cname, (_, fields) in structs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for f in fields:
            lines.append(f'  printf("{cname} {f} %zu\\n", offsetof({cname}, {f}));')
    lines += ['  printf("USF_MAX_MLP %d\\n", USF_MAX_MLP);', "  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "layout"
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    seen = {}
    for line in out.splitlines():
        parts = line.split()
        if parts[0] == "USF_MAX_MLP":
            assert int(parts[1]) == _lib.USF_MAX_MLP
            continue
        seen[(parts[0], parts[1])] = int(parts[2])
    for cname, (ctype, fields) in structs.items():
        assert seen[(cname, "size")] == ctypes.sizeof(ctype), cname
        declared = [f for f, _ in ctype._fields_]
        assert declared == fields, (cname, declared)          # same fields, same order
        for f in fields:
            assert seen[(cname, f)] == getattr(ctype, f).offset, (cname, f)
