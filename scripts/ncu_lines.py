"""Reads an .ncu-rep (no GPU needed) and prints, for one kernel, the warp-stall samples per CUDA source line
(`ncu --page source --print-source cuda,sass`; needs -lineinfo at compile time and --import-source on at capture).

    python scripts/ncu_lines.py gpurun_out/prof.ncu-rep mlp_coupling [top_n]
"""
import csv
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                          "regex:" + pat, "--launch-skip", "0", "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[h]
    si = hdr.index("# Samples")
    stall_cols = [(i, c) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
    per, cur = {}, None
    for r in rows[h + 1:]:
        if not r:
            continue
        if r[0] != "":
            cur = (r[0], ",".join(r[1:len(r) - (len(hdr) - 4)])[:100] if len(r) > len(hdr) else r[1][:100])
            continue
        if len(r) < len(hdr) or not r[si].isdigit():
            continue
        d = per.setdefault(cur, {"n": 0})
        d["n"] += int(r[si])
        for i, c in stall_cols:
            if r[i].isdigit():
                d[c] = d.get(c, 0) + int(r[i])
    tot = sum(d["n"] for d in per.values())
    print(f"{rows[1][1][:100]}\ntotal samples {tot}")
    agg = {}
    for d in per.values():
        for k, v in d.items():
            if k != "n":
                agg[k] = agg.get(k, 0) + v
    print("by reason:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
    for (ln, src), d in sorted(per.items(), key=lambda kv: -kv[1]["n"])[:top]:
        st = sorted(((v, k[6:]) for k, v in d.items() if k != "n" and v), reverse=True)[:3]
        print(f"{d['n']:5d} {100.0 * d['n'] / max(tot, 1):5.1f}%  L{ln:>5s} {src:100s} {st}")


if __name__ == "__main__":
    main()
