"""CPU: the oracle restatement reproduces the golden vectors that the
reference's own in-tree classes produced (tests/golden/make_golden.py)."""
import os

import pytest
import torch

from _cases import (build_flow, flow_from_case, flow_from_r2_case, golden_names, golden_r2_names, load_golden,
                    load_golden_r2)


@pytest.mark.parametrize("name", golden_names())
def test_oracle_reproduces_golden(O, name):
    g = load_golden(name)
    flow = flow_from_case(O, g["case"], ).double()
    flow.load_state_dict(g["state_dict"])
    x = g["x"].clone().requires_grad_(True)
    lp = flow.log_prob(x)
    assert lp.shape == (g["case"]["B"],)
    torch.testing.assert_close(lp, g["log_prob"], rtol=1e-12, atol=1e-12)
    with torch.no_grad():
        z = flow.backward(x)
        torch.testing.assert_close(z, g["latent"], rtol=1e-12, atol=1e-12)
        xs = flow.latent_to_data(g["z_sample"])
        torch.testing.assert_close(xs, g["x_from_z"], rtol=1e-12, atol=1e-12)
    grads = torch.autograd.grad(-lp.mean(), [x] + list(flow.parameters()), allow_unused=True)
    torch.testing.assert_close(grads[0], g["grad_x"], rtol=1e-10, atol=1e-12)
    for (n, _), gr in zip(flow.named_parameters(), grads[1:]):
        ref = g["grad_params"][n]
        if ref is None:
            assert gr is None
        else:
            torch.testing.assert_close(gr, ref, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("name", golden_names())
def test_golden_is_a_density(O, name):
    """Oracle self-check independent of upstream: log_prob == base(z) + log|det dz/dx| by autograd."""
    g = load_golden(name)
    D = g["case"]["D"]
    if D > 8:
        pytest.skip("jacobian check on small D only")
    flow = flow_from_case(O, g["case"]).double()
    flow.load_state_dict(g["state_dict"])
    x0 = g["x"][0]
    J = torch.autograd.functional.jacobian(lambda v: flow.backward(v[None])[0], x0)
    z0 = flow.backward(x0[None])
    base = flow._event_base.log_prob(z0)[0]
    torch.testing.assert_close(base + torch.linalg.slogdet(J)[1], g["log_prob"][0], rtol=1e-9, atol=1e-9)


def test_usflow_logdet_is_data_independent(O):
    """'Uniformly scaling': additive coupling => total log-det constant across inputs."""
    g = load_golden("us_d8_k2_densenn_trainable_normal")
    flow = flow_from_case(O, g["case"]).double()
    flow.load_state_dict(g["state_dict"])
    x = torch.randn(16, 8, dtype=torch.float64)
    with torch.no_grad():
        z = flow.backward(x)
        ld = flow.log_prob(x) - flow._event_base.log_prob(z)
    assert float(ld.max() - ld.min()) < 1e-10


@pytest.mark.parametrize("name", golden_r2_names())
def test_oracle_reproduces_round2_golden(O, name):
    """Image-shaped events, softplus scale activation, same-shape conditioner output, conditional flow: the restatement
    against what the reference's own classes produced (tests/golden/make_golden_r2.py)."""
    g = load_golden_r2(name)
    flow = flow_from_r2_case(O, g["case"]).double()
    flow.load_state_dict(g["state_dict"])
    x = g["x"].clone().requires_grad_(True)
    ctx = g["context"]
    lp = flow.log_prob(x, ctx)
    assert lp.shape == (g["case"]["B"],)
    torch.testing.assert_close(lp, g["log_prob"], rtol=1e-12, atol=1e-12)
    with torch.no_grad():
        torch.testing.assert_close(flow.backward(x, ctx), g["latent"], rtol=1e-12, atol=1e-12)
        torch.testing.assert_close(flow.latent_to_data(g["z_sample"], ctx), g["x_from_z"], rtol=1e-12, atol=1e-12)
        if g["ladj_quirk"] is not None:
            torch.testing.assert_close(torch.as_tensor(flow.log_abs_det_jacobian(x)), g["ladj_quirk"], rtol=1e-12, atol=1e-12)
    grads = torch.autograd.grad(-lp.mean(), [x] + list(flow.parameters()), allow_unused=True)
    torch.testing.assert_close(grads[0], g["grad_x"], rtol=1e-10, atol=1e-12)
    for (n, _), gr in zip(flow.named_parameters(), grads[1:]):
        ref = g["grad_params"][n]
        if ref is None:
            assert gr is None
        else:
            torch.testing.assert_close(gr, ref, rtol=1e-10, atol=1e-12)


def test_image_shaped_oracle_flow_is_a_density(O):
    """Upstream-independent check of the image-shaped semantics the oracle fixes (per-pixel C x C block, log-det once
    per pixel, `(B,)` coupling log-det): log_prob == base(z) + log|det dz/dx| by autograd on a [2, 3, 3] event."""
    g = load_golden_r2("nonus_img_c3_4x4_k3_channel_laplace")
    torch.manual_seed(5)
    flow = build_flow(O, "NonUSFlow", [2, 3, 3], 2, ("conv", [4]), affine_conjugation=True).double()
    x = torch.randn(2, 2, 3, 3, dtype=torch.float64)
    J = torch.autograd.functional.jacobian(lambda v: flow.backward(v.reshape(1, 2, 3, 3)).reshape(-1), x[0].reshape(-1))
    with torch.no_grad():
        z = flow.backward(x[:1])
        want = flow._event_base.log_prob(z)[0] + torch.linalg.slogdet(J)[1]
        torch.testing.assert_close(flow.log_prob(x)[0], want, rtol=1e-9, atol=1e-9)
    assert g["log_prob"].shape == (4,)


def test_nonusflow_log_abs_det_jacobian_quirk(O):
    """`NonUSFlow.log_abs_det_jacobian(x)` (`nf4ad/flows.py:160-169`) never advances x through the stack: it equals the
    sum over layers of -ladj_layer(layer.backward(x), x) evaluated AT THE DATA POINT -- not the flow's true log-det."""
    g = load_golden("nonus_d8_k2_conj_hh1")
    flow = flow_from_case(O, g["case"]).double()
    flow.load_state_dict(g["state_dict"])
    x = g["x"]
    with torch.no_grad():
        got = flow.log_abs_det_jacobian(x)
        want = 0
        for layer in reversed(flow.layers):
            want = want - layer.log_abs_det_jacobian(layer.backward(x), x)
        torch.testing.assert_close(got, want)
        true = flow.log_prob(x) - flow._event_base.log_prob(flow.backward(x))
        assert float((got - true).abs().max()) > 1e-3          # the quirk is real for data-dependent couplings


@pytest.mark.skipif(not __import__("oracle").ref_available(), reason="oracle/_ref snapshot missing (python oracle/make_ref.py)")
def test_reference_classes_equal_the_restatement(O):
    """The reference's own unmodified classes (snapshot `oracle/_ref`) and `oracle/nf4ad_restated` give the same numbers
    on the same shim: what the CPU baseline of kind "reference" times is what the parity tests check against."""
    import oracle
    R = oracle.load_ref()
    assert type(R.NonUSFlow).__name__ == "type" and R.NonUSFlow.__module__ == "nf4ad.flows"
    for name in golden_names():
        g = load_golden(name)
        if g["case"]["kind"] != "NonUSFlow":
            continue
        fr = flow_from_case(R, g["case"]).double()
        fr.load_state_dict(g["state_dict"])
        with torch.no_grad():
            torch.testing.assert_close(fr.log_prob(g["x"]), g["log_prob"], rtol=1e-12, atol=1e-12)


@pytest.mark.skipif(not __import__("oracle").ref_available(), reason="oracle/_ref snapshot missing (python oracle/make_ref.py)")
def test_reference_own_tests_pass_on_the_oracle_shim():
    """the reference's unmodified flow tests (`tests/test_flows.py`,
    `tests/test_adbench_flow_wrapper.py`, 9 tests; snapshot `oracle/_ref`) run against the oracle's `src.usflows` / `pyro`
    shim on the CPU.  (`tests/test_reference_on_gpu.py` runs the same files over the B200 drop-in.)"""
    import subprocess
    import sys
    from oracle import make_ref
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = make_ref.unpack()
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(root, "oracle", "shim"), os.path.join(ref, "src")]),
               PYTHONDONTWRITEBYTECODE="1", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "-m", "not slow",
                        "-c", os.path.join(ref, "pytest.ini"), "--rootdir", ref,
                        os.path.join(ref, "tests", "test_flows.py"), os.path.join(ref, "tests", "test_adbench_flow_wrapper.py")],
                       capture_output=True, text=True, env=env, cwd=ref, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
