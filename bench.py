#!/usr/bin/env python
"""Headline benchmark: `Flow.log_prob` samples/s on the MNIST-shape USFlow stack (BASELINE.json
configs[1]: D=784 NonUSFlow, K=8 coupling blocks, conditioner MLP 784-256-256-1568, LU + Householder
affine conjugation, Normal base), synthetic data, batch-sharded over N GPUs (weak scaling).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# ---- workload: BASELINE.json configs[1] / SURVEY.md section 8(d) "C2 MNIST" --------------------------------
D, K_BLOCKS, HIDDEN = 784, 8, [256, 256]
ROWS_PER_GPU = 65536
ALG_FLOP_PER_SAMPLE = 26.84e6      # mask-pruned, conditioner counted once (SURVEY 8d / BASELINE.md section 4)
ALG_BYTES_PER_SAMPLE = 4 * D + 4
LAST_LAYER_GAIN = 0.25             # trained-flow-like activations (see DESIGN.md "Synthetic weights")
CPU_SAMPLE_ROWS = 4096
EXEC_FLOP_PER_SAMPLE = 17.7e6      # what the launch chain executes: 9 merged dense maps + 8 pruned MLPs
# DRAM bytes per tensor-core launch at 65536 rows from the committed ncu capture: 9 GEMM launches at 174.4 MB
# (106.2 read + 68.2 written) and 8 fused conditioner launches at 120.3 MB (106.0 + 14.3), profiles/r1b_final
NCU_DRAM_BYTES_PER_LAUNCH = (9 * 174.4e6 + 8 * 120.3e6) / 17


def build_flow(ns, device, dtype=torch.float32):
    from _cases import build_flow as _bf, tame
    torch.manual_seed(0)
    flow = _bf(ns, "NonUSFlow", D, K_BLOCKS, ("mlp", HIDDEN), base="normal",
               affine_conjugation=True, prior_scale=1.0)
    tame(flow, LAST_LAYER_GAIN)
    return flow.to(device).eval()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index, self.err = [], None, index, ""
        self.sm = self.rows

    def _cmd(self, loop):
        cmd = ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"]
        return cmd + (["-lms", "50"] if loop else [])

    def start(self):
        try:
            self.proc = subprocess.Popen(self._cmd(True), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception as e:
            self.proc, self.err = None, repr(e)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def sample_once(self):
        """One synchronous sample (taken while kernels are in flight) in case the loop produced none."""
        try:
            r = subprocess.run(self._cmd(False), capture_output=True, text=True, timeout=20)
            if r.stdout.strip():
                self.rows.append(r.stdout.strip().splitlines()[0])
            else:
                self.err = (r.stderr or "").strip()[:200]
        except Exception as e:
            self.err = repr(e)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable: " + self.err]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.err = self.err or (self.proc.stderr.read() or "").strip()[:200]
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm)}
        if not sm and self.err:
            out["error"] = self.err
        return out


class NvmlSampler:
    """In-process NVML sampling thread (same counters as the nvidia-smi query above: SM clock, max SM
    clock, clock-event reasons).  Preferred over spawning `nvidia-smi -lms`, which was measured to stall
    kernel launches on this box (a 2 ms step became 12 ms while it polled).  NVML queries and CUDA launches share
    a driver lock: with three queries every 50 ms some runs lost 30-50 % of the timed region to stalled graph
    launches (host enqueue 1.2-2.6 ms per step instead of 0.04), so the thread asks for two values every 100 ms."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index, period_s=0.1):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except Exception:
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        self.period, self.sm, self.mask, self.stop_flag, self.rows = period_s, [], 0, False, []
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))

    def _one(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))

    def _loop(self):
        while not self.stop_flag:
            try:
                self._one()
            except Exception as e:          # keep the bench alive; report below
                self.err = repr(e)
                return
            time.sleep(self.period)

    def start(self):
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def sample_once(self):
        self._one()

    def stop(self):
        self.stop_flag = True
        if hasattr(self, "t"):
            self.t.join(timeout=2)
        try:        # one power reading, AFTER the timed region: every NVML query takes the driver lock the launches need
            self.rows.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
        except Exception:
            pass
        sm = sorted(self.sm)
        pw = sorted(self.rows)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for b, n in self.REASONS.items() if self.mask & b), "samples": len(sm),
                "power_w": pw[len(pw) // 2] if pw else None, "source": "nvml"}


class NullSampler:
    sm = []

    def start(self):
        pass

    def sample_once(self):
        pass

    def stop(self):
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampling disabled (USF_BENCH_SAMPLER=0)"], "samples": 0}


def make_sampler(index):
    if os.environ.get("USF_BENCH_SAMPLER", "1") == "0":
        return NullSampler()
    try:
        return NvmlSampler(index)
    except Exception:
        return ClockSampler(index)


def measured_peak_tflops():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
        except Exception:
            pass
    return 1400.0, "fallback (B200_PROFILING.md sustained)"


def cpu_reference_run(steps, warmup, rows):
    """The reference's CPU PyTorch path for the same workload: the oracle's restatement of
    `nf4ad.flows.NonUSFlow` + `MaskedAffineCoupling` on the `src.usflows` shim (the genuine USFlows /
    pyro packages are not installable), fp32, eval() + no_grad() as `adbench_wrapper.py:422-424`,
    all host threads."""
    import oracle
    O = oracle.load()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    flow = build_flow(O, "cpu")
    x = torch.randn(rows, D, generator=torch.Generator().manual_seed(42))
    with torch.no_grad():
        for _ in range(warmup):
            flow.log_prob(x)
        ts = []
        for _ in range(steps):
            t0 = time.perf_counter()
            flow.log_prob(x)
            ts.append(time.perf_counter() - t0)
    total = sum(ts)
    return rows * steps / total, cores, total / steps * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--rows", type=int, default=ROWS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--prewarm-s", type=float, default=1.0, help="untimed clock pre-warm (0 for profiler runs)")
    ap.add_argument("--train-steps", type=int, default=10, help="timed training steps (0 disables the leg)")
    ap.add_argument("--train-rows", type=int, default=4096, help="per-GPU training batch")
    ap.add_argument("--no-sweep", dest="sweep", action="store_false", help="skip the batch-size sweep")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "C2 MNIST-shape NonUSFlow log_prob scoring: D=784, K=8, MLP 784-256-256-1568, "
                          "LU+Householder conjugation, Normal base",
              "rows_per_gpu": args.rows, "global_rows": args.rows * max(world, 1), "parallelism": f"batch-shard x{world}",
              "l2_policy": "inputs larger than L2 (205 MB fp32 per GPU), no flush needed",
              "weights": "seeded init (seed 0), last conditioner layer x0.25"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        rows = CPU_SAMPLE_ROWS
        v, cores, ms = cpu_reference_run(args.steps, args.warmup, rows)
        line = {"impl": "reference", "metric": "log_prob samples/sec", "value": v, "unit": "samples/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": dict(config, rows_per_step=rows),
                "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                                 "sample": f"{rows} rows x {args.steps} steps of the same stack (oracle port of the "
                                           "reference's PyTorch path; real USFlows/pyro not installable)"},
                "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm (B200)
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import nf4ad_b200
    from nf4ad_b200 import _lib
    from nf4ad_b200.parallel import ShardedScorer
    P = nf4ad_b200.namespace()
    flow = build_flow(P, dev)
    flow.precision = args.precision
    B = args.rows
    gen = torch.Generator().manual_seed(42 + rank)
    x_host = torch.randn(B, D, generator=gen).pin_memory()
    x = x_host.to(dev)
    scorer = ShardedScorer(flow)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: K steps of the fused launch chain -----------------------------------
    with torch.no_grad():
        # pre-warm: a step is only ~2-3 ms, so run the same step for >= 1 s first to bring the GPU out of its
        # idle P-state to steady clocks (untimed), then the W warm-up steps the contract asks for
        # (untimed) until the step time has settled: a fresh process can sit in a slow state for a second or more
        # (idle clocks / power state; steps of 3-13 ms were seen), which must not leak into the timed region.  Batches
        # of 20 steps are timed with events until two consecutive batches agree within 3 %, for at least
        # `prewarm_s` and at most 8 s.
        t_pw = time.perf_counter()
        prev_ms = None
        while args.prewarm_s > 0:
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            for _ in range(20):
                lp = scorer.score_local(x)
            p1.record()
            torch.cuda.synchronize()
            cur_ms = p0.elapsed_time(p1)
            elapsed = time.perf_counter() - t_pw
            settled = prev_ms is not None and abs(cur_ms - prev_ms) <= 0.03 * prev_ms
            if (elapsed >= args.prewarm_s and settled) or elapsed > 8.0:
                break
            prev_ms = cur_ms
        for _ in range(args.warmup):
            lp = scorer.score_local(x)
        barrier()
        sampler = make_sampler(local)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        t_host = time.perf_counter()
        for i in range(args.steps):
            lp = scorer.score_local(x)
            if rank == 0 and i == args.steps // 2 and not sampler.sm:
                sampler.sample_once()          # kernels are in flight: launches are asynchronous
        host_enqueue_ms = (time.perf_counter() - t_host) * 1e3 / args.steps   # host time to enqueue one step
        e1.record()
        barrier()
        ms_total = e0.elapsed_time(e1)
        gstats = (ctypes.c_longlong * 4)()
        gfail = ctypes.create_string_buffer(160)
        _lib.lib().usf_debug_graph_stats(gstats, gfail, 160)
        graph_info = {"replays": int(gstats[0]), "captures": int(gstats[1]), "failed_captures": int(gstats[2]),
                      "eager_runs": int(gstats[3]), "last_failure": gfail.value.decode()}
        clocks = sampler.stop() if rank == 0 else None
        launches_per_step = flow.last_launches
        assert launches_per_step > 0, "fused CUDA path did not run"
        assert bool(torch.isfinite(lp).all())

        # ---- end to end through the public API: pinned host rows -> H2D -> log_prob -> scores D2H, every step
        # (the call sequence of ADBenchFlow.predict_score, adbench_wrapper.py:419-433) + score gather to rank 0
        for _ in range(12):        # untimed: covers the scorer's trial calls over its fp32-head candidates (parallel.py)
            scorer.predict_score_host(x_host)
        barrier()
        t0 = time.perf_counter()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            scores = scorer.predict_score_host(x_host, gather=True)
        f1.record()
        barrier()
        e2e_ms = max(f0.elapsed_time(f1), (time.perf_counter() - t0) * 1e3 if world == 1 else 0.0)

        # ---- per-launch device time of every kernel of one step (CUDA events on the launching stream)
        n_prof = launches_per_step * 3
        _lib.check(_lib.lib().usf_profile_begin(n_prof + 8))
        for _ in range(3):
            scorer.score_local(x)
        ms = (ctypes.c_float * (n_prof + 8))()
        tags = (ctypes.c_int * (n_prof + 8))()
        n = ctypes.c_int(0)
        _lib.check(_lib.lib().usf_profile_end(ms, tags, ctypes.byref(n)))
        per_tag = {}
        for i in range(n.value):
            per_tag.setdefault(tags[i], []).append(ms[i])
    # ---- batch-size sweep of the resident-input scoring call (1 024 ... 1 048 576 plus the reference's
    # own small batches); rank 0 only, short, reported as extra information
    sweep = {}
    if rank == 0 and args.sweep:
        with torch.no_grad():
            for rows_s in (64, 1024, 16384, 262144):
                for prec in ("bf16", "tf32x3", "fp32"):
                    if (prec == "fp32" and rows_s > 16384) or (prec == "tf32x3" and rows_s < 16384):
                        continue
                    flow.precision = prec
                    xs = torch.randn(rows_s, D, device=dev)
                    for _ in range(5):
                        flow.log_prob(xs)
                    torch.cuda.synchronize()
                    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    reps = 20 if rows_s <= 16384 else 5
                    h0.record()
                    for _ in range(reps):
                        flow.log_prob(xs)
                    h1.record()
                    torch.cuda.synchronize()
                    ms_s = h0.elapsed_time(h1) / reps
                    sweep[f"{prec}_rows{rows_s}"] = {"ms_per_call": ms_s, "samples_per_s": rows_s / (ms_s * 1e-3)}
            flow.precision = args.precision

    # ---- training step (second BASELINE metric): fwd + hand-written backward kernels + (DP all-reduce) + Adam
    train_ms, train_B, train_steps, train_graph = float("nan"), args.train_rows, args.train_steps, False
    train32_ms = train3_ms = float("nan")
    if train_steps > 0:
        from nf4ad_b200.parallel import DataParallelTrainer
        from nf4ad_b200.optim import FusedAdam
        xb = x[:train_B]
        for prec in ("bf16", "tf32x3", "fp32"):
            # "bf16" = mixed precision: bf16 tensor-core GEMMs with fp32 accumulation, fp32 parameters / gradients /
            # Adam state, LU layers applied through their per-step dense inverse; "tf32x3" = the same step with
            # 3xTF32 tensor-core GEMMs (fp32-grade); "fp32" = the all-fp32 kernels
            tflow = build_flow(P, dev).train()
            tflow.precision = prec
            opt = FusedAdam(tflow.parameters(), lr=1e-4)   # torch.optim.Adam's update via usf_adam_step; the step replays as one CUDA graph
            trainer = DataParallelTrainer(tflow, opt)
            trainer.broadcast_parameters()
            for _ in range(5):          # 3 eager steps + capture + first replay (single rank), all untimed
                trainer.step(xb)
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(train_steps):
                loss = trainer.step(xb)
            g1.record()
            barrier()
            if prec == "bf16":
                train_ms = g0.elapsed_time(g1)
                train_graph = trainer.graph_replays > 0
            elif prec == "tf32x3":
                train3_ms = g0.elapsed_time(g1)
            else:
                train32_ms = g0.elapsed_time(g1)
            assert bool(torch.isfinite(loss))
            del tflow, opt, trainer
    t = torch.tensor([ms_total, e2e_ms, train_ms, train32_ms, train3_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, train_ms, train32_ms, train3_ms = (float(v) for v in t)

    if rank == 0:
        n_gemm = sum(len(v) for k, v in per_tag.items() if k != 0)
        gemm_ms = sum(sum(v) for k, v in per_tag.items() if k != 0)
        steps_prof = 3
        gemm_ms_per_step = gemm_ms / steps_prof
        avg_launch_ms = gemm_ms / max(n_gemm, 1)
        flop_per_launch = ALG_FLOP_PER_SAMPLE * B / (n_gemm / steps_prof)
        achieved = flop_per_launch / (avg_launch_ms * 1e-3) / 1e12
        peak, peak_src = measured_peak_tflops()
        value = B * world * args.steps / (ms_total * 1e-3)
        e2e_value = B * world * args.steps / (e2e_ms * 1e-3)
        names = {0: "pack_input", 1: "affine_gemm", 2: "mlp_hidden_gemm", 3: "mlp_last_gemm+coupling", 4: "final_gemm+base",
                 5: "fused_conditioner+coupling"}
        line = {
            "metric": "log_prob samples/sec", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic", "config": config,
            "clocks": clocks, "host_enqueue_ms_per_step": host_enqueue_ms, "graph": graph_info,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(scorer.last_h2d_bytes),
                    "d2h_bytes_per_step": B * 4, "ms_per_step": e2e_ms / args.steps,
                    "host_input_bytes_per_step": B * D * 4,
                    "host_narrowing": bool(scorer.host_bf16 and args.precision == "bf16"),
                    "host_threads": int(scorer.host_threads),
                    "fp32_head_rows": (scorer._tune.get((B, D), {}).get("best", None) if scorer.host_bf16 else None)},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "tensor", "kernel": _lib.lib().usf_gemm_kernel_name(
                             _lib.USF_PREC_BF16 if args.precision == "bf16" else _lib.USF_PREC_FP32).decode(),
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                         "traffic_source": "ncu --set full dram__bytes_read+write per launch, mean over the 9 GEMM + 8 fused "
                                           "conditioner launches of a step, profiles/r1b_final/tc_kernels_full_raw.csv",
                         "executed_flop_per_sample": EXEC_FLOP_PER_SAMPLE,
                         "frac_executed": achieved / peak * EXEC_FLOP_PER_SAMPLE / ALG_FLOP_PER_SAMPLE,
                         "alg_flop_per_launch": flop_per_launch, "avg_launch_ms": avg_launch_ms,
                         "gemm_share_of_step": gemm_ms_per_step / (ms_total / args.steps),
                         "launch_ms_by_kind": {names[k]: sum(v) / len(v) for k, v in sorted(per_tag.items())},
                         "launches_by_kind": {names[k]: len(v) // steps_prof for k, v in sorted(per_tag.items())}},
        }
        if sweep:
            line["sweep"] = sweep
        if train_steps > 0:
            line["train"] = {"metric": "train samples/sec (fwd + bwd + Adam; bf16 tensor-core GEMMs, fp32 accumulate / "
                                       "parameters / optimizer state)", "graph_replay": bool(train_graph),
                             "value": train_B * world * train_steps / (train_ms * 1e-3), "unit": "samples/s",
                             "batch_per_gpu": train_B, "steps": train_steps, "ms_per_step": train_ms / train_steps,
                             "tf32x3_path": {"value": train_B * world * train_steps / (train3_ms * 1e-3),
                                             "ms_per_step": train3_ms / train_steps},
                             "fp32_path": {"value": train_B * world * train_steps / (train32_ms * 1e-3),
                                           "ms_per_step": train32_ms / train_steps}}
        if world == 1 and not args.no_cpu_baseline:
            v, cores, cms = cpu_reference_run(3, 1, CPU_SAMPLE_ROWS)
            line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                                    "sample": f"{CPU_SAMPLE_ROWS} rows x 3 steps of the same stack, fp32, "
                                              f"{cores} threads ({cms:.0f} ms/step)"}
            # the same leg also checks the timed GPU path against the fp64 oracle on the first 256 rows
            import oracle
            fo = build_flow(oracle.load(), "cpu").double()
            with torch.no_grad():
                ref = fo.log_prob(x_host[:256].double())
                got = flow.log_prob(x[:256]).double().cpu()
            line["cpu_baseline"]["gpu_vs_oracle_fp64_max_rel_err"] = float(((got - ref).abs() / ref.abs().clamp_min(1.0)).max())
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
