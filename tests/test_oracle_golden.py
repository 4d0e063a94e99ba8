"""CPU: the oracle restatement reproduces the golden vectors that the
reference's own in-tree classes produced (tests/golden/make_golden.py)."""
import os

import pytest
import torch

from _cases import flow_from_case, golden_names, load_golden


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


@pytest.mark.skipif(not os.path.isdir("/root/reference/tests"), reason="reference checkout only exists in the build container")
def test_reference_own_tests_pass_on_the_oracle_shim():
    """the reference's unmodified flow tests (`tests/test_flows.py`,
    `tests/test_adbench_flow_wrapper.py`, 9 tests) run against the oracle's `src.usflows` / `pyro` shim."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(root, "oracle", "shim"), "/root/reference/src"]),
               PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "-m", "not slow",
                        "/root/reference/tests/test_flows.py", "/root/reference/tests/test_adbench_flow_wrapper.py"],
                       capture_output=True, text=True, env=env, cwd="/tmp", timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
