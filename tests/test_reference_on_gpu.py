"""GPU: the reference's OWN, unmodified classes and tests on the B200 drop-in (parity definition 1).

The files come from the `oracle/_ref` snapshot (recipe `oracle/make_ref.py`; `/root/reference` does not exist on the
GPU box): `nf4ad/flows.py`, `nf4ad/transforms.py`, `nf4ad/adbench_wrapper.py` and the reference's
`tests/test_flows.py`, `tests/test_adbench_flow_wrapper.py` + `conftest.py`.  `src.usflows.*` and `pyro.*` resolve to
`nf4ad_b200/dropin`, i.e. to the CUDA kernels behind the C ABI.
"""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import oracle
from _cases import tame

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref snapshot missing (python oracle/make_ref.py)")]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_own_tests_pass_unmodified_on_the_b200_dropin():
    """`/root/reference/tests/test_flows.py:14-63` and `tests/test_adbench_flow_wrapper.py:56-158` (8 tests + 1 `slow`; `device`
    fixture -> cuda): NonUSFlow construction, sample, log_prob, ADBenchFlow fit / predict_score / predict / thresholds."""
    from oracle import make_ref
    import nf4ad_b200
    ref = make_ref.unpack()
    path = [os.path.join(nf4ad_b200.DROPIN, "_pyro"), nf4ad_b200.DROPIN, os.path.join(ref, "src"), ROOT,
            os.path.join(ROOT, "tests", "_plugins")]
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(path), PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-s", "-p", "no:cacheprovider", "-p", "usf_native_report",
                        "-m", "not slow", "-c", os.path.join(ref, "pytest.ini"), "--rootdir", ref,
                        os.path.join(ref, "tests", "test_flows.py"), os.path.join(ref, "tests", "test_adbench_flow_wrapper.py")],
                       capture_output=True, text=True, env=env, cwd=ref, timeout=900)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-4000:]
    m = re.search(r"(\d+) passed", out)
    assert m and int(m.group(1)) >= 8 and not re.search(r"\b\d+ (failed|error)", out), out[-2000:]      # 9 collected, the `slow` one deselected
    rep = re.search(r"USF_NATIVE_REPORT loaded=(\d) .*eager_runs=(\d+)", out)
    assert rep and rep.group(1) == "1", "the native library was not loaded by the reference's tests: " + out[-800:]
    assert int(rep.group(2)) > 0, "no fused launch chain ran during the reference's tests"


@pytest.fixture(scope="module")
def ref_on_cuda():
    """The reference's modules imported over the drop-in, in this process (the oracle swaps its own shim in and out
    around its imports, `oracle.activated`)."""
    from oracle import make_ref
    import nf4ad_b200
    nf4ad_b200.install_dropin(with_pyro=True)
    src = os.path.join(make_ref.unpack(), "src")
    sys.path.insert(0, src)
    for k in [k for k in sys.modules if k.split(".")[0] == "nf4ad"]:
        del sys.modules[k]
    try:
        import nf4ad.flows as rflows
        import nf4ad.adbench_wrapper as rwrap
        import pyro.distributions as pdist
        assert rflows.Flow.__module__ == "nf4ad_b200.flows", "src.usflows did not resolve to the drop-in"
        yield rflows, rwrap, pdist
    finally:
        sys.path.remove(src)


class SimpleConditioner(torch.nn.Module):            # tests/conftest.py:111-121
    def __init__(self, in_dim, out_dim, hidden=128):
        super().__init__()
        self.net = torch.nn.Sequential(torch.nn.Linear(in_dim, hidden), torch.nn.ReLU(), torch.nn.Linear(hidden, out_dim))

    def forward(self, x):
        return self.net(x)


def test_reference_classes_on_cuda_match_the_cpu_reference(ref_on_cuda):
    """Same seeded weights in (a) the reference's NonUSFlow over the drop-in on cuda and (b) the reference's NonUSFlow
    over the CPU oracle shim in fp64: log_prob, latents, samples agree to the fp32 tier, and the reference's unmodified
    `ADBenchFlow.predict_score` gives the same scores on both -- and the same as `nf4ad_b200.adbench.ADBenchFlow`."""
    rflows, rwrap, pdist = ref_on_cuda
    R = oracle.load_ref()
    D = 20
    dev = torch.device("cuda")
    torch.manual_seed(0)
    args = dict(in_dims=[D], coupling_blocks=3, prior_scale=1.0, affine_conjugation=True, conditioner_cls=SimpleConditioner,
                conditioner_args={"in_dim": D, "out_dim": 2 * D}, nonlinearity=torch.nn.ReLU())
    gpu = rflows.NonUSFlow(base_distribution=pdist.Normal(torch.zeros(D).to(dev), torch.ones(D).to(dev)), device=dev, **args)
    assert type(gpu).__module__ == "nf4ad.flows" and type(gpu.layers[1]).__module__ == "nf4ad.transforms"
    assert next(gpu.parameters()).is_cuda
    tame(gpu, 0.25)
    cpu = R.NonUSFlow(base_distribution=R.dist.Normal(torch.zeros(D), torch.ones(D)), device="cpu", **args)
    cpu.load_state_dict({k: v.cpu() for k, v in gpu.state_dict().items()})
    cpu64 = R.NonUSFlow(base_distribution=R.dist.Normal(torch.zeros(D, dtype=torch.float64), torch.ones(D, dtype=torch.float64)),
                        device="cpu", **args).double()
    cpu64.load_state_dict({k: v.cpu().double() for k, v in gpu.state_dict().items()})
    rng = np.random.RandomState(42)
    X = np.vstack([rng.randn(50, D), rng.randn(50, D) * 3 + 5]).astype(np.float32)
    x = torch.from_numpy(X)
    with torch.no_grad():
        gpu.eval()
        lp = gpu.log_prob(x.to(dev))
        assert gpu.last_launches > 0, "the reference's coupling class was not recognised by the fused path"
        ref = cpu64.log_prob(x.double())
        err = ((lp.double().cpu() - ref).abs() / ref.abs().clamp_min(1.0)).max()
        assert float(err) < 1e-4, float(err)
        z = gpu.backward(x.to(dev)).double().cpu()
        z_ref = cpu64.backward(x.double())          # the outlier rows reach |z| ~ 1e4: error relative to the row's scale
        assert float(((z - z_ref).abs().amax(1) / z_ref.abs().amax(1).clamp_min(1.0)).max()) < 1e-3
        s = gpu.sample([7])
        assert s.shape == (7, D) and s.is_cuda
    # the reference's wrapper, unmodified, around both
    wg = rwrap.ADBenchFlow(flow_model=gpu, epochs=2, batch_size=32, lr=1e-3, device="cuda", verbose=False)
    wc = R.ADBenchFlow(flow_model=cpu, epochs=2, batch_size=32, lr=1e-3, device="cpu", verbose=False)
    sg, sc = wg.predict_score(X), wc.predict_score(X)
    assert sg.shape == sc.shape == (100,)
    assert float(np.max(np.abs(sg - sc) / np.maximum(np.abs(sc), 1.0))) < 1e-4
    from nf4ad_b200.adbench import ADBenchFlow
    ours = ADBenchFlow(flow_model=gpu, epochs=2, batch_size=32, lr=1e-3, device="cuda", verbose=False).predict_score(X)
    assert float(np.max(np.abs(ours - sc) / np.maximum(np.abs(sc), 1.0))) < 1e-4
    # training through the reference's own loop (adbench_wrapper.py:347-404) on the CUDA autograd path, then scoring
    # again: finite, spread, separates the outliers, and the fused path follows the new weights
    Xtr = rng.randn(200, D).astype(np.float32)
    wg.fit(Xtr)
    s2 = wg.predict_score(X)
    assert np.all(np.isfinite(s2)) and s2.std() > 0 and np.median(s2[50:]) > np.median(s2[:50])
    assert float(np.abs(s2 - sg).mean()) > 1e-3
    assert len(wg.training_losses) == 2 and wg.training_losses[-1] < wg.training_losses[0]
