"""GPU: the mixed-precision training step's weight-space machinery (round 2) -- the structured LU inverse, affine runs
composed in weight space against the layer-wise pass, the packed `[s | t]` coupling gradient, weight gradients that
trail the backward chain, the library's own prioritised streams."""
import pytest
import torch

from _cases import build_flow, randomize_constants, tame

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    assert torch.cuda.is_available()
    import nf4ad_b200
    return nf4ad_b200.namespace()


@pytest.mark.parametrize("D", [16, 40, 64, 100, 784, 1000])
def test_lu_inverse_against_fp64(D):
    """usf_lu_inverse (`LUTransform.backward` of the tensor-core training tiers: both triangular inverses in one launch,
    identity right-hand sides, then their product) against the fp64 inverse of `(tril(L,-1)+I) triu(U)`
    (`USFlows LUTransform`, call sites nf4ad/flows.py:85,110), and against the row-wise solve it replaces."""
    from nf4ad_b200 import ops
    g = torch.Generator().manual_seed(D)
    L = (torch.randn(D, D, generator=g) * 0.05).cuda()
    U = (torch.randn(D, D, generator=g) * 0.05 + 1.5 * torch.eye(D)).cuda()
    A = ops.lu_inverse(L, U)
    eye = torch.eye(D, device="cuda")
    ref = torch.linalg.inv(((torch.tril(L, -1) + eye).double() @ torch.triu(U).double()))
    err = float((A.double() - ref).abs().max() / ref.abs().max())
    # D >= 256 and a multiple of 16 takes the 3xTF32 product (1e-5 grade), the rest the fp32 FFMA product
    assert err <= (2e-5 if (D % 16 == 0 and D >= 256) else 2e-6), err
    old = ops.lu_solve(eye, L, U, None, transpose=True)
    assert float((A - old).abs().max() / old.abs().max()) <= 3e-5


def test_lu_inverse_factors_only():
    """`usf_lu_inverse(A = NULL)` leaves U^{-1} and L^{-1} in the scratch buffer for the caller to multiply."""
    import ctypes as C
    from nf4ad_b200 import _lib
    D = 96
    g = torch.Generator().manual_seed(1)
    L = (torch.randn(D, D, generator=g) * 0.1).cuda()
    U = (torch.randn(D, D, generator=g) * 0.1 + torch.eye(D)).cuda()
    n = int(_lib.lib().usf_lu_inverse_scratch_floats(D))
    assert n >= 2 * D * D
    scratch = torch.zeros(n, device="cuda")
    _lib.check(_lib.lib().usf_lu_inverse(_lib.ptr(L), _lib.ptr(U), D, None, _lib.ptr(scratch), _lib.stream()), "usf_lu_inverse")
    eye = torch.eye(D, device="cuda")
    Z, W = scratch[:D * D].view(D, D), scratch[D * D:2 * D * D].view(D, D)
    assert float((Z.double() - torch.linalg.inv(torch.triu(U).double())).abs().max()) < 1e-5
    assert float((W.double() - torch.linalg.inv((torch.tril(L, -1) + eye).double())).abs().max()) < 1e-5
    # structure: U^{-1} upper, L^{-1} unit lower -- the skipped blocks are exact zeros
    assert float(torch.tril(Z, -1).abs().max()) == 0.0 and float(torch.triu(W, 1).abs().max()) == 0.0
    with pytest.raises(_lib.USFError):
        _lib.check(_lib.lib().usf_lu_inverse(None, _lib.ptr(U), D, None, _lib.ptr(scratch), _lib.stream()), "usf_lu_inverse")


def _grads(flow, x, **attrs):
    for k, v in attrs.items():
        setattr(flow, k, v)
    flow.zero_grad(set_to_none=True)
    loss = -flow.log_prob(x).mean()
    loss.backward()
    torch.cuda.synchronize()
    return float(loss.detach()), {n: p.grad.detach().clone() for n, p in flow.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("D,K,hidden,B", [(64, 3, [128, 128], 256), (784, 2, [256], 1024)])
def test_composed_affine_runs_match_the_layerwise_pass(O, P, D, K, hidden, B):
    """`Flow._compose_affine_runs` (one GEMM per affine run on the batch; the run's matrix / shift / log-det evaluated in
    weight space on side streams) against the layer-wise pass of the same tier and against the fp64 oracle's loss and
    gradients (the reference's arithmetic: nf4ad/flows.py:160-169 + adbench_wrapper.py:383-391).  Every gradient of the
    3xTF32 tier within 1e-2 of fp64 in norm, every gradient of the bf16 tier with the right direction and norm, and the
    composed pass never further from fp64 than 1.5x the layer-wise pass of its tier."""
    torch.manual_seed(0)
    fo = build_flow(O, "NonUSFlow", D, K, ("mlp", hidden), affine_conjugation=True, prior_scale=1.0)
    tame(fo, 0.25)
    randomize_constants(fo, 3)
    flow = build_flow(P, "NonUSFlow", D, K, ("mlp", hidden), affine_conjugation=True, prior_scale=1.0)
    flow.load_state_dict(fo.state_dict())
    flow = flow.to("cuda").train()
    xc = torch.randn(B, D, generator=torch.Generator().manual_seed(5))
    x = xc.cuda()
    fo = fo.double()
    l64 = -fo.log_prob(xc.double()).mean()
    l64.backward()
    g64 = {n: p.grad.detach() for n, p in fo.named_parameters() if p.grad is not None}
    l64 = float(l64.detach())

    def dist(g):
        out = {}
        for n, ref in g64.items():
            if float(ref.norm()) > 1e-9:
                out[n] = float((g[n].double().cpu() - ref).norm() / ref.norm())
        return out

    l32, g32 = _grads(flow, x, precision="fp32")
    assert set(g32) == set(g64) and abs(l32 - l64) <= 1e-5 * max(1.0, abs(l64))
    e32 = dist(g32)
    assert max(e32.values()) < 1e-2, max(e32.values())
    for prec, ltol in (("tf32x3", 1e-4), ("bf16", 1e-2)):
        errs = {}
        for which, composed in (("composed", True), ("layer-wise", False)):
            loss, g = _grads(flow, x, precision=prec, compose_affine=composed)
            assert abs(loss - l64) <= ltol * max(1.0, abs(l64)), (prec, which, loss, l64)
            assert set(g) == set(g64)
            errs[which] = dist(g)
            for n, e in errs[which].items():
                assert torch.isfinite(g[n]).all(), (prec, which, n)
                if prec == "tf32x3":
                    # the tier's gradients sit 3e-3 .. 5e-3 from fp64 at D = 784 in either pass (the tensor cores'
                    # fp32 accumulation over K = 784 .. 1024; the fp32 FFMA kernels: 7e-5)
                    assert e <= 1e-2, (prec, which, n, e, e32[n])
                else:
                    ref, got = g64[n], g[n].double().cpu()
                    cos = float((ref * got).sum() / (ref.norm() * got.norm()).clamp_min(1e-30))
                    assert cos > 0.98 and 0.9 < float(got.norm() / ref.norm()) < 1.1, (prec, which, n, cos)
        # composing in weight space does not cost accuracy: no gradient further from fp64 than 1.5x the layer-wise one
        for n, e in errs["composed"].items():
            assert e <= 1.5 * errs["layer-wise"][n] + (1e-3 if prec == "tf32x3" else 2e-2), (prec, n, e, errs["layer-wise"][n])
    flow.compose_affine = True


def test_composed_pass_is_used_and_cleans_up(P):
    """The composed pass really is the one that runs (a `LinearTCFn` with a transposed weight per run) and leaves no
    single-use state behind on the modules."""
    from nf4ad_b200 import ops, transforms
    torch.manual_seed(0)
    flow = build_flow(P, "NonUSFlow", 64, 3, ("mlp", [64]), affine_conjugation=True).to("cuda").train()
    flow.precision = "bf16"
    x = torch.randn(64, 64, device="cuda")
    seen = []
    orig = ops.linear_fn

    def spy(x, W, bias, relu=False, w_transposed=False, operands=None):
        seen.append((bool(w_transposed), operands is not None))
        return orig(x, W, bias, relu, w_transposed=w_transposed, operands=operands)

    ops.linear_fn = spy
    try:
        (-flow.log_prob(x).mean()).backward()
    finally:
        ops.linear_fn = orig
    assert sum(1 for t, _ in seen if t) == 4                # K + 1 affine runs, one GEMM each on the batch
    assert any(pre for _, pre in seen)                      # conditioner operands prepared ahead
    assert transforms._COMPOSE is None
    for m in flow.modules():
        assert "_A_pre" not in m.__dict__ and "_usf_pre" not in m.__dict__


def test_packed_coupling_gradient_equals_the_split_form():
    """`CouplingPackedFn` (one `[s | t]` tensor in, one gradient tensor out) is `CouplingFn` on the two halves, bit for
    bit -- both directions, both scale activations (`MaskedAffineCoupling._parse_params`, nf4ad/transforms.py:52-55)."""
    from nf4ad_b200 import ops
    g = torch.Generator().manual_seed(2)
    B, D = 37, 24
    mask = (torch.arange(D) % 2).float().cuda()
    for inverse in (False, True):
        for act in (0, 1):
            x = torch.randn(B, D, generator=g).cuda().requires_grad_()
            st = (0.3 * torch.randn(B, 2 * D, generator=g)).cuda().requires_grad_()
            x2, st2 = x.detach().clone().requires_grad_(), st.detach().clone().requires_grad_()
            y, l = ops.CouplingPackedFn.apply(x, st, mask, 2.0, inverse, act)
            y2, l2 = ops.CouplingFn.apply(x2, st2[:, :D], st2[:, D:], mask, 2.0, inverse, act)
            w, wl = torch.randn(B, D, generator=g).cuda(), torch.randn(B, generator=g).cuda()
            ((y * w).sum() + (l * wl).sum()).backward()
            ((y2 * w).sum() + (l2 * wl).sum()).backward()
            assert torch.equal(y, y2) and torch.equal(l, l2)
            assert torch.equal(x.grad, x2.grad) and torch.equal(st.grad, st2.grad)


def test_trailing_weight_gradients_equal_the_joined_form():
    """`ops._WgradSide`: the weight-gradient GEMM left running on the partner stream gives the very gradients of the
    form that joins before returning -- including a weight that is itself the result of an op on another stream."""
    from nf4ad_b200 import ops
    g = torch.Generator().manual_seed(4)
    B, K, N = 2048, 256, 128
    x0 = torch.randn(B, K, generator=g).cuda()
    W0 = (0.1 * torch.randn(N, K, generator=g)).cuda()
    m = (torch.arange(K) % 2).float().cuda()
    out = {}
    for defer in (True, False):
        ops._WGRAD_DEFER = defer
        try:
            x, W = x0.clone().requires_grad_(), W0.clone().requires_grad_()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                Wm = W * m                                   # (how the mask-folded first conditioner layer arrives)
            torch.cuda.current_stream().wait_stream(side)
            with ops.tc_training(1):
                y = ops.linear_fn(x, Wm, None, True)
                z = ops.linear_fn(y, (0.1 * torch.ones(64, N, device="cuda")).requires_grad_(), None, False)
            z.square().sum().backward()
            torch.cuda.synchronize()
            out[defer] = (x.grad.clone(), W.grad.clone())
        finally:
            ops._WGRAD_DEFER = True
    assert torch.equal(out[True][0], out[False][0]) and torch.equal(out[True][1], out[False][1])
    assert float(out[True][1][:, 0::2].abs().max()) == 0.0     # masked columns get no gradient


def test_own_streams_are_distinct_and_prioritised():
    """`usf_stream_create`: 40 streams, 40 different handles (torch's pool would repeat after 32), usable as torch
    streams, priorities clamped to the device's range."""
    from nf4ad_b200 import _lib
    streams = [_lib.new_stream("cuda:0", priority=-(i % 9)) for i in range(40)]
    assert len({s.cuda_stream for s in streams}) == 40
    x = torch.ones(1024, device="cuda")
    torch.cuda.synchronize()
    for s in streams:
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            x.add_(1.0) if s is streams[0] else None
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    assert float(x[0]) == 2.0
    for s in streams:
        _lib.check(_lib.lib().usf_stream_destroy(_lib.C.c_void_p(s.cuda_stream)), "usf_stream_destroy")


def test_conditional_flow_trains_in_the_tensor_core_tier(O, P):
    """A conditional flow (soft training: `ConditionalDenseNN` conditioners on `cat([context, x])`, nf4ad/flows.py:56-57,
    transforms.py:71-74) in the bf16 training tier: the coupling mask is folded into the first conditioner layer with
    the context columns left alone, the affine runs are composed.  Loss within the tier of the fp64 oracle's, gradients
    in its direction, and `Flow.fit` runs on the graph trainer."""
    import numpy as np
    D, B = 32, 256
    torch.manual_seed(2)
    prior = lambda ns, dev: ns.dist.Uniform(torch.tensor(0.0, device=dev), torch.tensor(0.2, device=dev))
    kw = dict(affine_conjugation=True, prior_scale=1.0, soft_training=True)
    fo = build_flow(O, "NonUSFlow", D, 3, ("cond2", [64]), training_noise_prior=prior(O, "cpu"), **kw)
    tame(fo, 0.25)
    fp = build_flow(P, "NonUSFlow", D, 3, ("cond2", [64]), training_noise_prior=prior(P, "cpu"), **kw)
    fp.load_state_dict(fo.state_dict())
    fo, fp = fo.double(), fp.to("cuda").train()
    g = torch.Generator().manual_seed(3)
    x, ctx = torch.randn(B, D, generator=g), 0.2 * torch.rand(B, 1, generator=g)
    l64 = -fo.log_prob(x.double(), ctx.double()).mean()
    l64.backward()
    g64 = {n: p.grad for n, p in fo.named_parameters() if p.grad is not None}
    fp.precision = "bf16"
    loss = -fp.log_prob(x.cuda(), ctx.cuda()).mean()
    loss.backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(l64)) <= 1e-2 * max(1.0, abs(float(l64)))
    for n, p in fp.named_parameters():
        if n in g64 and float(g64[n].norm()) > 1e-6:
            got, ref = p.grad.double().cpu(), g64[n]
            cos = float((ref * got).sum() / (ref.norm() * got.norm()).clamp_min(1e-30))
            assert cos > 0.97, (n, cos)
    fp.zero_grad(set_to_none=True)
    data = torch.randn(512, D, generator=g) * 0.5 + 0.3
    # (a) While `loss` -- and with it the autograd graph of that step on the DEFAULT stream -- is alive, so are its
    # gradient accumulators; a captured step that reaches them would make the legacy stream wait for a capturing one.
    # The trainer must notice, warn, recover (streams joined, the generator out of its capture state: the soft-training
    # noise draws random numbers) and go on with eager steps.
    import warnings
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        losses = fp.fit(data, torch.optim.Adam, {"lr": 1e-3}, batch_size=64, epochs=2)
    assert len(losses) == 2 and np.all(np.isfinite(losses))
    if fp.fit_graph_replays == 0:
        assert any("capture of the training step failed" in str(w.message) for w in caught)
    # (b) the normal case: nothing of an earlier step is kept -> graph replay
    del loss
    import gc
    gc.collect()
    losses = fp.fit(data, torch.optim.Adam, {"lr": 1e-3}, batch_size=64, epochs=4)
    assert len(losses) == 4 and np.all(np.isfinite(losses)) and losses[-1] < losses[0]
    assert fp.fit_graph_replays >= 8 * 4 - 4


def test_many_flows_share_one_set_of_streams(P):
    """A sweep builds and trains many flows in one process (hyperopt): the side streams are per device, not per flow."""
    from nf4ad_b200 import _lib
    x = torch.randn(64, 32, device="cuda")
    seen = None
    for i in range(4):
        torch.manual_seed(i)
        flow = build_flow(P, "NonUSFlow", 32, 2, ("mlp", [32]), affine_conjugation=True).to("cuda").train()
        flow.precision = "bf16"
        (-flow.log_prob(x).mean()).backward()
        torch.cuda.synchronize()
        n = len(_lib._STREAM_POOL)
        assert seen is None or n == seen, (seen, n)
        seen = n
