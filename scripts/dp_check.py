"""Data-parallel gradient check on N GPUs (torchrun): the gradients `DataParallelTrainer` leaves in the flat buffer after
backward + exchange, against the single-process gradients of the full batch.  Prints the worst relative distance (the
bf16 tier rounds each rank's composed-run gradient before the weight-space chain: a few 1e-2 on the LU factors) and
checks that every rank holds the same numbers."""
import os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nf4ad_b200
from _cases import build_flow, tame
from nf4ad_b200.parallel import DataParallelTrainer
from nf4ad_b200.optim import FusedAdam

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
P = nf4ad_b200.namespace()
D, K, hidden, B = 784, 3, [256, 256], 1024 * world


def make():
    torch.manual_seed(0)
    f = build_flow(P, "NonUSFlow", D, K, ("mlp", hidden), affine_conjugation=True, prior_scale=1.0)
    tame(f, 0.25)
    f = f.to("cuda").train()
    f.precision = "bf16"
    return f


x = torch.randn(B, D, generator=torch.Generator().manual_seed(5)).cuda()
ref = make()
(-ref.log_prob(x).mean()).backward()
gref = {n: p.grad.detach().clone() for n, p in ref.named_parameters()}
for _ in (0,):
    flow = make()
    tr = DataParallelTrainer(flow, FusedAdam(flow.parameters(), lr=0.0))
    shard = x[rank * (B // world):(rank + 1) * (B // world)]
    tr._zero_grad()
    loss = tr.loss_fn(shard)
    loss.backward()
    tr.allreduce_gradients()
    torch.cuda.synchronize()
    worst, name = 0.0, None
    for n, p in flow.named_parameters():
        e = float((p.grad - gref[n]).norm() / gref[n].norm().clamp_min(1e-30))
        if e > worst:
            worst, name = e, n
    sent = sum((b[1] - b[0]) * 4 for b in tr._buckets)
    # every rank must hold the same gradients
    flat = tr._flat.clone()
    dist.all_reduce(flat, op=dist.ReduceOp.MAX)
    spread = float((flat - tr._flat).abs().max())
    if rank == 0:
        print(f"worst gradient distance from the single-process full batch {worst:.3e} ({name}); "
              f"bucket bytes sent {sent / 1e6:.1f} MB in {len(tr._buckets)} buckets, max spread between ranks {spread:.3e}")
    tr.close()
dist.barrier()
dist.destroy_process_group()
