"""CPU, world_size 2 over gloo: the sharding / score-gather / gradient all-reduce plumbing of
nf4ad_b200.parallel.  The scoring callable here is the CPU oracle flow (test infrastructure): the
product kernels themselves need a GPU and are covered by the -m gpu tests."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_rows_in_order():
    from nf4ad_b200.parallel import shard_bounds
    for n in (0, 1, 7, 8, 1000, 65536 + 3):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from _cases import build_flow, tame
        from nf4ad_b200.parallel import DataParallelTrainer, ShardedScorer
        O = oracle.load()
        torch.manual_seed(0)
        flow = build_flow(O, "NonUSFlow", 6, 2, ("mlp", [8]), affine_conjugation=True)
        tame(flow, 0.5)
        X = torch.randn(101, 6, generator=torch.Generator().manual_seed(1))
        # ---- sharded scoring == single-process scoring
        scorer = ShardedScorer(flow, score_fn=lambda x: flow.log_prob(x))
        scores = scorer.predict_score(X)
        with torch.no_grad():
            ref = -flow.log_prob(X)
        assert scores.shape == (101,)
        assert torch.allclose(scores, ref, rtol=1e-6, atol=1e-6)
        # ---- data-parallel step == single-process step on the concatenated batch
        import copy
        single = copy.deepcopy(flow)
        opt_s = torch.optim.SGD(single.parameters(), lr=1e-5)
        opt_s.zero_grad()
        (-single.log_prob(X[:64]).mean()).backward()
        opt_s.step()
        opt = torch.optim.SGD(flow.parameters(), lr=1e-5)
        os.environ["USF_DP_BUCKET_MB"] = "0.0002"          # ~50 floats per bucket: many buckets, sent from the hooks
        tr = DataParallelTrainer(flow, opt)
        assert len(tr._buckets) > 3 and tr._buckets[0][0] == 0 and tr._buckets[-1][1] == tr._flat.numel()
        assert all(a[1] == b[0] for a, b in zip(tr._buckets, tr._buckets[1:]))
        tr.broadcast_parameters()
        lo, hi = (0, 32) if rank == 0 else (32, 64)
        tr.step(X[lo:hi])
        assert all(tr._sent) and not tr._syncing
        for (n, a), (_, b) in zip(flow.named_parameters(), single.named_parameters()):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), n
        # every .grad is a view into the one flat buffer (nothing to flatten / copy back), and a second step with the
        # hooks switched off (all buckets sent after backward) gives the same update as the single process
        base = tr._flat.data_ptr()
        off = 0
        for p in tr.params:
            assert p.grad.data_ptr() == base + 4 * off and p.grad.shape == p.shape
            off += p.numel()
        tr.close()
        os.environ["USF_DP_OVERLAP"] = "0"
        tr2 = DataParallelTrainer(flow, opt)
        assert not tr2._hooks
        opt_s.zero_grad()
        (-single.log_prob(X[:64]).mean()).backward()
        opt_s.step()
        tr2.step(X[lo:hi])
        for (n, a), (_, b) in zip(flow.named_parameters(), single.named_parameters()):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), (n, float((a - b).abs().max()), float(b.abs().max()))
        # rank-sharded batches of `fit`: both ranks' shards of one global batch make the single-process batch
        from nf4ad_b200.parallel import rank_batches, shared_permutation
        perm = shared_permutation(101, "cpu", torch.Generator().manual_seed(5 + rank))     # rank 0's draw wins
        mine = list(rank_batches(perm, 32, rank, world))
        gathered = [None, None]
        dist.all_gather_object(gathered, [b.tolist() for b in mine])
        for step in range(len(mine)):
            glob = gathered[0][step] + gathered[1][step]
            assert glob == perm[step * 32: step * 32 + len(glob)].tolist()
        if rank == 0:
            open(tmp, "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_world2_scoring_and_dp_step(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    marker = str(tmp_path / "ok")
    mp.spawn(_worker, args=(2, port, marker), nprocs=2, join=True)
    assert open(marker).read() == "ok"


def test_stage_plan_covers_rows_in_order():
    """Host logic of the end-to-end scoring pipeline (`parallel.stage_plan`): chunks are contiguous, in order and cover
    every row exactly once; only a leading block of at most `raw_rows` rows is sent unstaged and never so much that no
    full chunk is left to stage under it; no chunk exceeds the ring's chunk size; the schedule ends on a short chunk."""
    from nf4ad_b200.parallel import stage_plan
    for chunk in (1, 7, 16384):
        for raw in (0, chunk // 2, chunk, 3 * chunk):
            for n in (1, chunk - 1, chunk, chunk + 1, 2 * chunk, 3 * chunk + 5, 4 * chunk, 10 * chunk + chunk // 2 + 1):
                if n <= 0:
                    continue
                plan = stage_plan(n, chunk, raw)
                assert plan[0][0] == 0 and plan[-1][1] == n
                assert all(a[1] == b[0] for a, b in zip(plan, plan[1:]))
                assert all(0 < hi - lo <= chunk for lo, hi, _ in plan)
                flags = [st for _, _, st in plan]
                assert flags == sorted(flags)                      # unstaged head first, then staged chunks only
                head = sum(hi - lo for lo, hi, st in plan if not st)
                assert head == max(0, min(raw, n - chunk))
                assert plan[-1][2] and plan[-1][1] - plan[-1][0] <= chunk // 2 + 1       # short copy + compute tail
    assert stage_plan(65536, 16384, 16384) == [(0, 16384, False), (16384, 32768, True), (32768, 49152, True),
                                               (49152, 57344, True), (57344, 65536, True)]
