"""Turns the ncu captures of scripts/gpu_ncu_r2.sh (read here, no GPU) into the tracked summaries under profiles/r2/:
   tc_kernels_full_raw.csv   ncu --page raw of the --set full capture (the two tensor-core kernels)
   tc_kernels_summary.txt    the counters DESIGN.md quotes, per captured launch
   stall_lines_*.txt         warp-stall samples per CUDA source line (scripts/ncu_lines.py)
   ncu_dram_bytes.json       DRAM bytes per launch and kind: cold (default cache control) and warm (--cache-control none,
                             with / without the serpentine tile order) -- bench.py reads `bytes_per_launch` from it
   launches_steady.csv       the launch list of two steady steps
"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles", "r2")
KIND = (("mlp_coupling", "conditioner+coupling"), ("tc_gemm", "affine_gemm"), ("convert_rows", "pack_input"))


def kind_of(name):
    for pat, k in KIND:
        if pat in name:
            return k
    return "other"


def metric_rows(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 8]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rows[1:]:
        try:
            key = (int(r[ix["ID"]]), r[ix["Kernel Name"]])
            v, unit = float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]]
        except (ValueError, KeyError, IndexError):
            continue
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ns": 1e-9, "ms": 1e-3, "s": 1.0}.get(unit, 1.0)
        per.setdefault(key, {})[r[ix["Metric Name"]]] = v * scale
    return per


def dram_by_kind(path):
    out = {}
    for (_, name), m in metric_rows(path).items():
        k = kind_of(name)
        b = m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        out.setdefault(k, []).append((b, m.get("dram__bytes_read.sum", 0.0), m.get("dram__bytes_write.sum", 0.0),
                                      m.get("gpu__time_duration.sum", 0.0), m.get("lts__t_sector_hit_rate.pct", 0.0)))
    return {k: {"launches": len(v), "bytes": sum(x[0] for x in v) / len(v), "read": sum(x[1] for x in v) / len(v),
                "write": sum(x[2] for x in v) / len(v), "time_us": 1e6 * sum(x[3] for x in v) / len(v),
                "l2_hit_pct": sum(x[4] for x in v) / len(v)} for k, v in out.items()}


def main():
    os.makedirs(DST, exist_ok=True)
    rep = os.path.join(SRC, "prof_r2.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(os.path.join(DST, "tc_kernels_full_raw.csv"), "w").write(raw)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
    lines, cold = ["ncu --set full --clock-control none (default cache control: cold caches per replay), one line set per captured launch", ""], {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        lines.append(name[:110])
        for w in want:
            for i, h in enumerate(hdr):
                if h.startswith(w):
                    lines.append(f"    {h:78s} {r[i]:>16s} {units[i]}")
                    break
        lines.append("")
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", "")), units[hdr.index("dram__bytes_read.sum")]
            wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", "")), units[hdr.index("dram__bytes_write.sum")]
            sc = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}
            cold.setdefault(kind_of(name), []).append(rd[0] * sc[rd[1]] + wr[0] * sc[wr[1]])
        except (ValueError, KeyError):
            pass
    open(os.path.join(DST, "tc_kernels_summary.txt"), "w").write("\n".join(lines) + "\n")
    for pat in ("mlp_coupling", "tc_gemm"):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_lines.py"), rep, pat, "40"], capture_output=True, text=True).stdout
        open(os.path.join(DST, f"stall_lines_{pat}.txt"), "w").write(out)
    warm = dram_by_kind(os.path.join(SRC, "ncu_warm_dram.csv"))
    noserp = dram_by_kind(os.path.join(SRC, "ncu_warm_dram_noserp.csv")) if os.path.exists(os.path.join(SRC, "ncu_warm_dram_noserp.csv")) else {}
    cold = {k: sum(v) / len(v) for k, v in cold.items()}
    per_step = lambda d: sum(v["bytes"] * v["launches"] for v in d.values())
    summary = {
        "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, one 18-launch step of the C2 chain at 65536 rows "
                  "(scripts/gpu_ncu_r2.sh); `bytes_per_launch` = warm caches (--cache-control none, serpentine tile order), the "
                  "state the chain really runs in; `cold_bytes_per_launch` = default ncu cache control (flushed per replay)",
        "bytes_per_launch": {k: v["bytes"] for k, v in warm.items()},
        "warm_serpentine": warm, "warm_first_to_last_order": noserp, "cold_bytes_per_launch": cold,
        "warm_bytes_per_step": per_step(warm), "warm_bytes_per_step_first_to_last_order": per_step(noserp) if noserp else None,
    }
    summary["bytes_per_launch"]["final_gemm+base"] = summary["bytes_per_launch"].get("affine_gemm")
    json.dump(summary, open(os.path.join(DST, "ncu_dram_bytes.json"), "w"), indent=1)
    shutil.copyfile(os.path.join(SRC, "launches_steady.csv"), os.path.join(DST, "launches_steady.csv"))
    print(json.dumps({k: summary[k] for k in ("bytes_per_launch", "cold_bytes_per_launch", "warm_bytes_per_step",
                                              "warm_bytes_per_step_first_to_last_order")}, indent=1))


if __name__ == "__main__":
    main()
