"""Multi-GPU plumbing (one process per GPU, torch.distributed / NCCL over NVLink).

Scoring shards by rows: every op of the density path is per-row, weights are
replicated, so there is NO collective on the data path -- only one all-gather of the (rows,) fp32
scores.  Training is data parallel: every `.grad` is a view into one flat fp32 buffer that NCCL averages in place, in
buckets handed over while the backward pass is still running, then the same optimizer step on every rank -- all inside the
step's CUDA graph (`DataParallelTrainer`).
"""
import os
import time

import torch
import torch.distributed as dist


def shard_bounds(n_rows, rank, world):
    """Contiguous row block of `rank`: sizes differ by at most one row, order preserved."""
    base, rem = divmod(n_rows, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def rank_batches(perm, batch_size, rank, world):
    """The reference's mini-batch schedule (`adbench_wrapper.py:364-377`: consecutive `batch_size` slices of one shuffled
    index vector) under data parallelism: every rank walks the SAME global batches of the same `perm` and takes its
    `shard_bounds` slice of each, so one optimizer step still sees exactly `batch_size` samples and an epoch sees every
    sample once.  A trailing batch with fewer rows than ranks is dropped on every rank (it cannot be sharded)."""
    n = perm.shape[0]
    for i in range(0, n, batch_size):
        idx = perm[i:i + batch_size]
        if world > 1:
            if idx.shape[0] < world:
                return
            lo, hi = shard_bounds(idx.shape[0], rank, world)
            idx = idx[lo:hi]
        yield idx


def shared_permutation(n, device, generator, shuffle=True, group=None):
    """One shuffled index vector, identical on every rank (rank 0's draw is broadcast)."""
    perm = torch.randperm(n, device=device, generator=generator) if shuffle else torch.arange(n, device=device)
    rank, world = _world(group)
    if world > 1:
        dist.broadcast(perm, 0, group=group)
    return perm


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(index):
    """NUMA node of the host memory closest to GPU `index` (sysfs), or None when the platform does not say."""
    try:
        p = torch.cuda.get_device_properties(index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_numa(index, local_rank=0, local_world=1):
    """Multi-rank hosts: every rank streams hundreds of MB per scoring call out of host DRAM (205 MB of fp32 rows per 65536 x
    784 call), and a rank whose pinned staging buffers / input pages sit on the other socket pulls them across the
    inter-socket link -- with 8 ranks on one host that, not PCIe, set the end-to-end rate in round 1 (21 GB/s per GPU at
    N = 8, and N = 4 slower per GPU than N = 8).  Before it allocates any host buffer a rank therefore pins itself to the CPUs
    of its GPU's NUMA node (first-touch then places its pages there), sharing them with the other local ranks of that node.
    Returns a dict describing what was done (for the bench record); a no-op where sysfs does not expose the topology or
    USF_NUMA_BIND=0."""
    global _NUMA_DONE
    if _NUMA_DONE is not None:          # once per process: a second call would split the already split CPU set again
        return _NUMA_DONE
    _NUMA_DONE = _bind_to_gpu_numa(index, local_rank, local_world)
    return _NUMA_DONE


_NUMA_DONE = None


def _bind_to_gpu_numa(index, local_rank, local_world):
    info = {"bound": False}
    if os.environ.get("USF_NUMA_BIND", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return info
    node = gpu_numa_node(index)
    if node is None:
        return info
    try:
        node_cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        allowed = os.sched_getaffinity(0) & node_cpus
        if len(allowed) < 2:
            return dict(info, node=node, reason="fewer than 2 allowed CPUs on the GPU's node")
        # the local ranks whose GPUs hang off the same node share its CPUs evenly
        peers = [r for r in range(local_world) if gpu_numa_node(r) == node] or [local_rank]
        mine = sorted(allowed)
        if len(peers) > 1 and len(mine) >= 2 * len(peers):
            k = peers.index(local_rank) if local_rank in peers else 0
            per = len(mine) // len(peers)
            mine = mine[k * per:(k + 1) * per]
        os.sched_setaffinity(0, set(mine))
        torch.set_num_threads(max(1, len(mine)))
        return {"bound": True, "node": node, "cpus": len(mine), "peers_on_node": len(peers)}
    except Exception as e:            # never fail a scoring call over a placement hint
        return dict(info, node=node, reason=repr(e))


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def stage_plan(n_rows, chunk_rows, raw_rows):
    """Row chunks of one host -> device scoring call as (lo, hi, staged) triples, in order and covering [0, n_rows):
    the leading `raw_rows` (at most n_rows - chunk_rows; 0 for pageable input) go over the link as they are (fp32), so
    that it is busy from t = 0 while the host stages the next chunk; every later chunk is staged through the pinned ring
    (narrowed or copied); the last full-size chunk is cut in two halves (short copy + compute tail)."""
    half = max(1, chunk_rows // 2)
    plan, lo = [], 0
    raw = max(0, min(raw_rows, n_rows - chunk_rows))
    while lo < raw:
        plan.append((lo, min(raw, lo + chunk_rows), False))
        lo = plan[-1][1]
    while lo < n_rows:
        hi = min(n_rows, lo + chunk_rows)
        if n_rows - lo <= chunk_rows and n_rows - lo > half:
            hi = lo + half
        plan.append((lo, hi, True))
        lo = hi
    return plan


class ShardedScorer:
    """Batch-sharded anomaly scoring: `predict_score` semantics of the reference wrapper
    (`/root/reference/src/nf4ad/adbench_wrapper.py:406-433`, score = -log_prob) over N GPUs."""

    def __init__(self, flow, group=None, score_fn=None):
        self.flow, self.group = flow, group
        self.score_fn = score_fn or (lambda x: flow.log_prob(x))
        self.rank, self.world = _world(group)
        self.chunk_rows = int(os.environ.get("USF_HOST_CHUNK_ROWS", "16384"))   # rows per H2D / compute pipeline stage
        self._copy_stream = None
        # bf16 tier: narrow the rows to bf16 on the host cores before the PCIe copy (half the bytes; the tier rounds its
        # input to bf16 as its first device step, so the scores are bit-identical).  Only with the default score function.
        # Measured (scripts/host_narrow.py): one rank with the host's 16 threads narrows at 103 GB/s of
        # fp32, twice the PCIe rate, and the call gets 1.35x faster; with 8 threads the conversion is the bottleneck, and
        # two ranks with 12 threads each contend for host DRAM (5.3 vs 4.3 ms) -- so it is on by default only for a single
        # rank per host with >= 16 host threads (USF_HOST_BF16=1/0 forces it).
        local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
        self.numa = {"bound": False}
        if local_world > 1 and torch.cuda.is_available():
            # before any pinned buffer exists: this rank's host pages belong next to its GPU (bind_to_gpu_numa)
            self.numa = bind_to_gpu_numa(torch.cuda.current_device(), int(os.environ.get("LOCAL_RANK", "0")), local_world)
        try:
            avail = len(os.sched_getaffinity(0))
        except AttributeError:
            avail = os.cpu_count() or 1
        # after the binding the affinity mask IS this rank's share of the host
        share = avail if self.numa.get("bound") else max(1, avail // local_world)
        self.host_threads = int(os.environ.get("USF_HOST_THREADS", "0")) or share
        want = os.environ.get("USF_HOST_BF16")
        self.host_bf16 = score_fn is None and (want == "1" or (want is None and local_world == 1 and self.host_threads >= 16))
        # Pageable rows (the numpy arrays the reference's callers pass): always staged through the pinned ring by the host
        # thread pool -- narrowed in the bf16 tier, copied otherwise -- so that the PCIe copy is asynchronous and pipelined
        # with the staging of the next chunk and the compute of the previous one (a pageable cudaMemcpy is neither):
        # 65536 x 784 rows, bf16 tier: 20.2 -> 3.9 ms per call (scripts/pageable_e2e.py).
        self.host_staging = os.environ.get("USF_HOST_STAGING", "1") != "0"
        self.narrow_ok = score_fn is None       # a custom score function may not take bf16 rows
        self.last_h2d_bytes = 0
        self.raw_rows = int(os.environ.get("USF_HOST_RAW_ROWS", "16384"))    # leading rows sent as fp32 (pinned input)
        self.autotune = "USF_HOST_RAW_ROWS" not in os.environ                # ... unless fixed: found per shape by trying
        self._tune = {}
        self._ring = None

    def score_local(self, x_dev):
        """log_prob of rows already resident on this rank's device."""
        return self.score_fn(x_dev)

    def gather(self, local, sizes=None):
        """All ranks' (rows_r,) vectors concatenated in rank order (returned on every rank)."""
        if self.world == 1:
            return local
        if sizes is None or len(set(sizes)) == 1:
            out = torch.empty(self.world * local.numel(), device=local.device, dtype=local.dtype)
            dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
            return out
        m = max(sizes)
        pad = torch.zeros(m, device=local.device, dtype=local.dtype)
        pad[:local.numel()] = local
        out = torch.empty(self.world * m, device=local.device, dtype=local.dtype)
        dist.all_gather_into_tensor(out, pad, group=self.group)
        return torch.cat([out[r * m:r * m + sizes[r]] for r in range(self.world)])

    def predict_score_host(self, x_host, gather=False):
        """This rank's rows from host memory (pinned or pageable) -> device -> log_prob -> -scores back on the host.
        H2D and D2H happen inside the call (adbench_wrapper.py:419,433)."""
        scores = self._scores_from_host(x_host)
        if gather:
            scores = self.gather(scores)
        return scores.cpu()

    def _scores_from_host(self, x_host):
        """-log_prob of host rows as a device vector: the copy / staging pipeline picked for this input."""
        dev = next(self.flow.parameters()).device if hasattr(self.flow, "parameters") else torch.device("cpu")
        n = x_host.shape[0]
        stageable = (dev.type == "cuda" and x_host.dtype == torch.float32 and x_host.dim() == 2
                     and x_host.stride(1) == 1 and n >= self.chunk_rows)
        bf16_tier = self.narrow_ok and getattr(self.flow, "precision", None) == "bf16"
        if bf16_tier and stageable and hasattr(self.flow, "resolve_precision"):
            # "bf16" is a verified tier (Flow._tier): rows may only be narrowed on the host when this stack really runs
            # its bf16 kernels -- a stack routed to the 3xTF32 kernels wants the fp32 rows
            tier = self.flow.cached_precision()
            if tier is None:
                with torch.no_grad():
                    tier = self.flow.resolve_precision(x_host[:256].to(dev))
            bf16_tier = tier == "bf16"
        pinned = x_host.is_pinned() if stageable else False
        # pinned rows: narrowing only where it was measured to pay (host_bf16); pageable rows: always through the ring
        narrow = stageable and bf16_tier and (self.host_bf16 if pinned else (self.host_staging or self.host_bf16))
        staged_copy = stageable and not narrow and not pinned and self.host_staging
        self.last_h2d_bytes = x_host.numel() * x_host.element_size()
        tune = None
        if narrow and self.autotune and pinned and n >= 3 * self.chunk_rows:
            # Host speed differs from box to box (measured 71-103 GB/s on the same pool), so the fp32 head that balances
            # conversion against the link is found by trying: the first calls with a given shape cycle through the
            # candidates twice -- every one is a full, valid scoring call -- then the fastest stays (None = plain copy).
            tune = self._tune.setdefault((n, x_host.shape[1]), {"i": 0, "t": {}, "best": False})
            if tune["best"] is False:
                cands = [self.raw_rows] + [c for c in (self.raw_rows * 3 // 2, self.raw_rows // 2, 2 * self.raw_rows, None)
                                           if c is None or 0 < c <= n - self.chunk_rows]
                choice = cands[tune["i"] % len(cands)]
                t_start = time.perf_counter()
            else:
                choice, tune = tune["best"], None
            narrow = choice is not None
        else:
            choice = self.raw_rows
        if narrow or staged_copy:
            scores = self._score_staged(x_host, dev, choice, narrow)
        elif dev.type != "cuda" or n < 2 * self.chunk_rows:
            x = x_host.to(dev, non_blocking=True)
            with torch.no_grad():
                scores = -self.score_fn(x)
        else:
            # chunked, double-buffered: the H2D copy of chunk i+1 (copy stream) overlaps the launch chain
            # of chunk i (current stream); rows are independent so results are identical
            cur = torch.cuda.current_stream(dev)
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(dev)
            cs = self._copy_stream
            cs.wait_stream(cur)
            scores = torch.empty(n, device=dev, dtype=torch.float32)
            pending = []
            for lo in range(0, n, self.chunk_rows):
                hi = min(n, lo + self.chunk_rows)
                with torch.cuda.stream(cs):
                    xc = x_host[lo:hi].to(dev, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(cs)
                pending.append((lo, hi, xc, ev))
            with torch.no_grad():
                for lo, hi, xc, ev in pending:
                    cur.wait_event(ev)
                    xc.record_stream(cur)
                    torch.neg(self.score_fn(xc), out=scores[lo:hi])
        if tune is not None:
            torch.cuda.current_stream(dev).synchronize()       # trial calls only: time the whole pipeline
            dt = time.perf_counter() - t_start
            tune["t"][choice] = min(dt, tune["t"].get(choice, dt))
            tune["i"] += 1
            if tune["i"] >= 2 * len(cands):
                tune["best"] = min(tune["t"], key=tune["t"].get)
        return scores

    def _score_staged(self, x_host, dev, raw_rows, narrow):
        """Three-stage pipeline over row chunks: host cores stage chunk i+1 into a pinned ring -- narrowed to bf16
        (usf_host_f32_to_bf16) or copied as fp32 (usf_host_copy_f32, pageable rows of the other tiers) -- while the copy
        stream moves chunk i over PCIe and the launch chain of chunk i-1 runs.
        The chunk schedule is `stage_plan` (fp32 head from pinned memory, short last chunk)."""
        from . import _lib
        n, D = x_host.shape
        cur = torch.cuda.current_stream(dev)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(dev)
        cs = self._copy_stream
        rdt = torch.bfloat16 if narrow else torch.float32
        if self._ring is None or self._ring[0][0].shape != (self.chunk_rows, D) or self._ring[0][0].dtype != rdt:
            self._ring = [[torch.empty(self.chunk_rows, D, dtype=rdt).pin_memory(), None] for _ in range(3)]
        stage = _lib.lib().usf_host_f32_to_bf16 if narrow else _lib.lib().usf_host_copy_f32
        plan = stage_plan(n, self.chunk_rows, raw_rows if x_host.is_pinned() else 0)
        cs.wait_stream(cur)
        scores = torch.empty(n, device=dev, dtype=torch.float32)
        k = 0
        self.last_h2d_bytes = sum((hi - lo) * D * (2 if (staged and narrow) else 4) for lo, hi, staged in plan)
        with torch.no_grad():
            for lo, hi, staged in plan:
                if staged:
                    slot = self._ring[k % len(self._ring)]
                    k += 1
                    if slot[1] is not None:
                        slot[1].synchronize()          # the copy that last read this staging buffer has finished
                    src = x_host[lo:hi]
                    _lib.check(stage(_lib.ptr(src), src.stride(0) if hi - lo > 1 else D, _lib.ptr(slot[0]), D, hi - lo, D,
                                     self.host_threads), "usf_host_f32_to_bf16 / usf_host_copy_f32")
                    buf = slot[0][:hi - lo]
                else:
                    slot, buf = None, x_host[lo:hi]
                with torch.cuda.stream(cs):
                    xc = buf.to(dev, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(cs)
                if slot is not None:
                    slot[1] = ev
                cur.wait_event(ev)
                xc.record_stream(cur)
                torch.neg(self.score_fn(xc), out=scores[lo:hi])
        return scores

    def predict_score(self, X_all):
        """Full (N, D) host array on every rank -> (N,) scores; rank r scores rows shard_bounds(N, r, W)."""
        X_all = torch.as_tensor(X_all, dtype=torch.float32)
        n = X_all.shape[0]
        lo, hi = shard_bounds(n, self.rank, self.world)
        local = self._scores_from_host(X_all[lo:hi])       # staged / pipelined like predict_score_host
        sizes = [shard_bounds(n, r, self.world)[1] - shard_bounds(n, r, self.world)[0] for r in range(self.world)]
        return self.gather(local, sizes).cpu()


def _is_capturing(st):
    with torch.cuda.stream(st):
        return torch.cuda.is_current_stream_capturing()


class DataParallelTrainer:
    """Data-parallel NLL training step (`adbench_wrapper.py:375-392`): per-rank micro-batch, gradients averaged over the
    ranks (loss is a per-rank mean), global-norm clipping after the reduction, identical optimizer step on every rank.

    Gradient exchange (world > 1).  Every parameter's `.grad` is a view into ONE persistent flat fp32 buffer laid out in
    parameter order, so there is nothing to flatten, scale or copy back: NCCL averages (`ReduceOp.AVG`) the buffer in place,
    in buckets of >= `USF_DP_BUCKET_MB` (default 8 MB).  The backward pass finishes the blocks of the stack in parameter
    order (log_prob walks the layers last -> first, its backward first -> last), so a bucket is handed to NCCL on a side
    stream the moment autograd has accumulated its last gradient (`register_post_accumulate_grad_hook`) and travels over
    NVLink while the remaining blocks' backward kernels run; the step joins the side stream before clipping / the
    optimizer.  All of it -- kernels, side-stream forks, collectives -- is captured in the step's CUDA graph.

    CUDA-graph replay.  At the reference's own batch sizes (32-64) the step is bound by the
    host: ~300 autograd nodes and kernel launches cost 5-16 ms per step while the kernels need a fraction of
    that.  After `graph_warmup` eager steps with a given batch shape the whole step -- forward, the hand-written
    backward kernels, clipping, the optimizer update -- is captured once into a CUDA graph and every later step
    with that shape is one `cudaGraphLaunch` on a static input buffer (multi-rank NCCL jobs capture the gradient
    all-reduce with it).  Used when the parameters live on a CUDA device and the optimizer was built with
    `capturable=True` (its step counters live on the device); anything else, and any shape seen fewer than
    `graph_warmup` times, runs eagerly."""

    def __init__(self, flow, optimizer, group=None, gradient_clip=None, loss_fn=None, use_graph=None,
                 graph_warmup=3):
        self.flow, self.opt, self.group, self.clip = flow, optimizer, group, gradient_clip
        self.loss_fn = loss_fn or (lambda batch: -flow.log_prob(batch).mean())
        self.rank, self.world = _world(group)
        self.params = [p for p in flow.parameters() if p.requires_grad]
        capturable = all(g.get("capturable", False) for g in getattr(optimizer, "param_groups", [])) \
            if getattr(optimizer, "param_groups", None) else False
        on_cuda = bool(self.params) and all(p.is_cuda for p in self.params)
        if use_graph is None:
            use_graph = os.environ.get("USF_TRAIN_GRAPH", "1") != "0"
        # multi-rank: the gradient all-reduce is captured with the step (NCCL collectives are graph-capturable); other
        # backends (gloo in the CPU tests) run eagerly
        nccl = self.world == 1 or (dist.is_available() and dist.is_initialized() and dist.get_backend(group) == "nccl")
        if self.world > 1 and os.environ.get("USF_TRAIN_GRAPH_MULTI", "1") == "0":
            nccl = False
        self.use_graph = bool(use_graph) and nccl and capturable and on_cuda
        self.graph_warmup = int(graph_warmup)
        self._capture_stream = None
        self._graphs = {}      # batch shape -> [seen, CUDAGraph | None, static input, static loss, hyper-parameter signature]
        self.graph_replays = 0
        self.graph_error = None     # the exception that switched graph replay off, if any
        self._flat = None
        self._hooks = []
        self.bucket_bytes = int(float(os.environ.get("USF_DP_BUCKET_MB", "8")) * (1 << 20))
        self.overlap = os.environ.get("USF_DP_OVERLAP", "1") != "0"
        if self.world > 1 and self.params:
            self._setup_flat_gradients()

    # ------------------------------------------------------------------ flat gradient buffer + bucketed exchange
    def _setup_flat_gradients(self):
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self._flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self._bucket_of, self._buckets = {}, []            # param id -> bucket; bucket = [start, end, n_params]
        off, start, count = 0, 0, 0
        for i, p in enumerate(self.params):
            n = p.numel()
            p.grad = self._flat[off:off + n].view_as(p)
            self._bucket_of[id(p)] = len(self._buckets)
            off += n
            count += 1
            if (off - start) * 4 >= self.bucket_bytes or i == len(self.params) - 1:
                self._buckets.append([start, off, count])
                start, count = off, 0
        self._pending = [b[2] for b in self._buckets]
        self._sent = [False] * len(self._buckets)
        self._syncing = False
        self._main_stream = None
        # (highest priority: a collective that queues for SMs behind the step's own kernels holds up the optimizer that
        # waits for it -- 8 GPUs: 4.08 -> 3.75 ms per step)
        self._comm_stream = torch.cuda.Stream(dev, priority=-5) if dev.type == "cuda" else None
        self._nccl = dev.type == "cuda" and dist.get_backend(self.group) == "nccl"
        if self.overlap:
            self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]

    def close(self):
        """Detach the gradient hooks (a trainer lives for one `fit`; the flow outlives it)."""
        for h in self._hooks:
            h.remove()
        self._hooks = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _on_grad(self, p):
        if not self._syncing:
            return
        b = self._bucket_of.get(id(p))
        if b is None:
            return
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._send_bucket(b)

    def _send_bucket(self, b):
        if self._sent[b]:
            return
        self._sent[b] = True
        start, end, _ = self._buckets[b]
        chunk = self._flat[start:end]
        if self._comm_stream is not None:
            # the bucket's gradients were produced on the step's main stream, on the stream autograd runs this hook on,
            # and -- LU inverses, weight gradients -- on the flow's side streams: wait for everything enqueued on any of
            # them so far (the rest of the backward pass keeps running beside the collective)
            from . import ops
            dev = chunk.device
            waits = {torch.cuda.current_stream(dev), self._main_stream}
            waits.update(getattr(self.flow, "_side_streams", None) or [])
            waits.update(st for (i, _), st in ops._WGRAD_STREAMS.items() if i == dev.index)
            # only streams in the same capture state as this hook's: a capturing stream must not wait for one outside
            # its capture (the weight-gradient partner of the eager warm-up steps' stream, say), and the other way round
            capturing = torch.cuda.is_current_stream_capturing()
            for st in waits:
                if st is not None and _is_capturing(st) == capturing:
                    self._comm_stream.wait_stream(st)
            with torch.cuda.stream(self._comm_stream):
                self._reduce(chunk)
        else:
            self._reduce(chunk)

    def _reduce(self, chunk):
        if self._nccl:
            dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=self.group)
        else:                                              # gloo (CPU tests): no AVG
            dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)
            chunk.div_(self.world)

    def _begin_sync(self):
        """Zero the flat buffer (autograd accumulates into the views in place) and arm the bucket counters."""
        self._flat.zero_()
        for p, (s0, e0) in zip(self.params, self._param_ranges()):
            if p.grad is None or p.grad.data_ptr() != self._flat.data_ptr() + 4 * s0:
                p.grad = self._flat[s0:e0].view_as(p)      # someone set it to None / replaced it: re-attach the view
        self._pending = [b[2] for b in self._buckets]
        self._sent = [False] * len(self._buckets)
        self._main_stream = torch.cuda.current_stream(self._flat.device) if self._flat.is_cuda else None
        self._syncing = True

    def _param_ranges(self):
        r = self.__dict__.get("_ranges")
        if r is None:
            r, off = [], 0
            for p in self.params:
                r.append((off, off + p.numel()))
                off += p.numel()
            self._ranges = r
        return r

    def _finish_sync(self):
        """Buckets whose parameters saw no gradient this step (or all of them, without the hooks) go out now, in order;
        then the step joins the side stream."""
        self._syncing = False
        for b in range(len(self._buckets)):
            self._send_bucket(b)
        if self._comm_stream is not None:
            torch.cuda.current_stream(self._flat.device).wait_stream(self._comm_stream)

    def broadcast_parameters(self, src=0):
        if self.world > 1:
            for p in self.flow.parameters():
                dist.broadcast(p.data, src, group=self.group)
            for b in self.flow.buffers():
                dist.broadcast(b.data, src, group=self.group)

    def allreduce_gradients(self):
        """Averages the gradients over the ranks (world > 1): sends whatever buckets the backward hooks have not sent yet
        and waits for all of them."""
        if self.world == 1 or self._flat is None:
            return
        self._finish_sync()

    def _clip(self):
        """Global-norm clipping (`adbench_wrapper.py:388-389`).  With `FusedAdam` the coefficient goes into the update
        kernel (a device scalar) instead of a scaling pass over the gradients."""
        from .optim import FusedAdam, SophiaG
        if not isinstance(self.opt, (FusedAdam, SophiaG)):
            torch.nn.utils.clip_grad_norm_(self.params, self.clip)
            return
        grads = [p.grad for p in self.params if p.grad is not None]
        if not grads:
            return
        total = torch.linalg.vector_norm(torch.stack(torch._foreach_norm(grads)))
        coef = torch.clamp(float(self.clip) / (total + 1e-6), max=1.0)
        if self.opt.clip_coef is None:
            self.opt.clip_coef = torch.ones((), device=grads[0].device, dtype=torch.float32)
        self.opt.clip_coef.copy_(coef)

    def _zero_grad(self):
        if self._flat is not None:
            self._begin_sync()
        else:
            self.opt.zero_grad(set_to_none=True)

    def _step_stream(self, device):
        from . import _lib
        if self._capture_stream is None or self._capture_stream.device != device:
            self._capture_stream = _lib.pooled_stream(device, "capture", priority=-8)
        return self._capture_stream

    def _eager_step(self, batch):
        if batch.is_cuda and self.use_graph:
            # The warm-up steps run on the stream the step will be captured on, not on the caller's: a parameter's
            # gradient accumulator keeps the stream of its first use for as long as anything references the graph, and
            # one left on the legacy default stream makes a later capture fail ("would make the legacy stream depend on
            # a capturing stream").
            st, cur = self._step_stream(batch.device), torch.cuda.current_stream(batch.device)
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                out = self._eager_step_on_current(batch)
            batch.record_stream(st)
            cur.wait_stream(st)
            return out
        return self._eager_step_on_current(batch)

    def _eager_step_on_current(self, batch):
        self._zero_grad()
        loss = self.loss_fn(batch)
        loss.backward()
        self.allreduce_gradients()
        if self.clip is not None:
            self._clip()
        self.opt.step()
        return loss.detach()

    def _capture(self, batch):
        static_x = batch.detach().clone()
        dot = os.environ.get("USF_GRAPH_DOT")        # debugging: dump the captured step's node / edge list (Graphviz)
        graph = torch.cuda.CUDAGraph(keep_graph=True) if dot else torch.cuda.CUDAGraph()
        if dot:
            graph.enable_debug_mode()
        if self._flat is None:
            self.opt.zero_grad(set_to_none=True)
        # The step is captured on a stream of the highest priority: graph kernel nodes inherit the priority of the stream
        # they were captured on, and the batch-sized chain (the step's critical path) must not queue behind the
        # weight-space chains the flow forks beside it (`Flow._compose_affine_runs`, lower priorities).
        # thread_local: NCCL's watchdog thread may query events while this thread captures
        with torch.cuda.graph(graph, stream=self._step_stream(static_x.device), capture_error_mode="thread_local"):
            if self._flat is not None:
                self._begin_sync()          # the memset of the flat gradient buffer is part of the replayed step
            loss = self.loss_fn(static_x)
            loss.backward()
            self.allreduce_gradients()
            if self.clip is not None:
                self._clip()
            self.opt.step()
            static_loss = loss.detach()
        if dot:
            graph.debug_dump(dot)
        return graph, static_x, static_loss

    def _hyper_sig(self):
        """Scalars a captured optimizer launch bakes in (lr, betas, eps, weight decay): a scheduler that changes them
        must trigger a re-capture, not be ignored by the replay."""
        sig = []
        for g in getattr(self.opt, "param_groups", []):
            sig.append(tuple((k, float(v) if isinstance(v, (int, float)) else tuple(v) if isinstance(v, (tuple, list)) else None)
                             for k, v in sorted(g.items()) if k in ("lr", "betas", "eps", "weight_decay")))
        return tuple(sig)

    def step(self, batch):
        from . import _lib
        try:
            return self._step(batch)
        finally:
            _lib.bump_weights_epoch()     # eager FusedAdam / graph replay: parameters changed behind torch's version counters

    def _step(self, batch):
        if not self.use_graph or not batch.is_cuda:
            return self._eager_step(batch)
        key = (tuple(batch.shape), batch.dtype, batch.device.index)
        slot = self._graphs.setdefault(key, [0, None, None, None, None])
        if slot[1] is not None and slot[4] != self._hyper_sig():
            slot[1] = None                                # lr / betas changed since the capture: record the step again
        if slot[1] is None:
            slot[0] += 1
            if slot[0] <= self.graph_warmup:          # real (eager) steps; they also initialise the optimizer state
                return self._eager_step(batch)
            try:
                slot[1], slot[2], slot[3] = self._capture(batch)     # records the step, does not run it
                slot[4] = self._hyper_sig()
            except Exception as e:                     # e.g. a conditioner that is not capture-safe
                import warnings
                self.use_graph = False
                self.graph_error = e.with_traceback(None)      # (no frames: they would keep the failed step's graph alive)
                warnings.warn(f"nf4ad_b200: CUDA-graph capture of the training step failed ({type(e).__name__}: {e}); "
                              "continuing with eager steps", RuntimeWarning)
                self._abandon_capture(batch.device)
                return self._eager_step(batch)
        slot[2].copy_(batch)
        slot[1].replay()
        self.graph_replays += 1
        return slot[3].clone()

    def _abandon_capture(self, device):
        """After a failed capture: drop matrices a half-recorded pass left on the LU layers, let every side stream the
        step forks (LU inversions, weight-gradient GEMMs) drain, and clear half-accumulated gradients."""
        from . import ops
        for m in self.flow.modules():
            m.__dict__.pop("_A_pre", None)
            m.__dict__.pop("_usf_pre", None)
        self.flow.__dict__.pop("_cond_pending", None)
        torch.cuda.synchronize(device)
        for st in list(getattr(self.flow, "_side_streams", None) or []) + list(ops._WGRAD_STREAMS.values()):
            st.synchronize()
        self.opt.zero_grad(set_to_none=True)
        # A capture that dies half way leaves the device's default generator in its "capturing" state (capture_end never
        # reached the generator's epilogue): the next random draw outside a graph -- the soft-training noise of the eager
        # fallback -- then raises "Offset increment outside graph capture".  An empty capture runs prologue and epilogue.
        try:
            with torch.cuda.device(device):
                with torch.cuda.graph(torch.cuda.CUDAGraph(), capture_error_mode="thread_local"):
                    pass
        except Exception:
            pass
