"""SASS opcode summary of the shipped library (runs on the build box, no GPU): per kernel, the counts of the Blackwell
tensor-core / TMEM / TMA mnemonics (B200_PROFILING.md "What proves a Blackwell-native kernel") and of the legacy paths
that must be absent.  Writes profiles/<round>/sass_opcodes.txt.

    python scripts/sass_summary.py [out_file]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "nf4ad_b200", "lib", "libusflow_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UTCBAR",
         "UTCCP", "SYNCS", "HMMA", "HGMMA", "LDGSTS", "MUFU", "ATOMG", "RED", "FFMA", "LDG", "STG"]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2", "sass_opcodes.txt")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            op, mods = m.group(1), m.group(2)
            cur["__total__"] += 1
            for w in WATCH:
                if op == w or (w in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTCBAR", "LDTM", "MUFU", "LDG", "STG") and op.startswith(w)):
                    cur[w] += 1
                    if w in ("UTCHMMA", "UTMALDG", "UTCBAR", "LDTM", "MUFU") and mods:
                        cur[w + mods] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    lines = ["SASS opcode summary of nf4ad_b200/lib/libusflow_b200.so (cuobjdump -sass, sm_100a)", ""]
    totals = collections.Counter()
    for (mangled, c), name in zip(per.items(), demangle):
        short = re.sub(r"\(.*", "", name)
        keys = [k for k in c if k != "__total__"]
        if not any(k.startswith(("UTC", "LDTM", "STTM", "UTMA", "HMMA", "HGMMA")) for k in keys):
            continue
        lines.append(f"{short}   ({c['__total__']} instructions)")
        for k in sorted(keys):
            if k.split(".")[0] in ("FFMA", "LDG", "STG", "RED", "SYNCS"):
                continue
            lines.append(f"    {k:<40s} {c[k]}")
        for k in keys:
            totals[k.split(".")[0]] += c[k] if "." not in k else 0
        lines.append("")
    lines.append("whole library (every kernel):")
    whole = collections.Counter()
    for c in per.values():
        for k, v in c.items():
            if "." not in k and k != "__total__":
                whole[k] += v
    for k in WATCH:
        lines.append(f"    {k:<12s} {whole.get(k, 0)}")
    lines.append("")
    lines.append("tcgen05.mma -> UTCHMMA(.2CTA), tcgen05.ld -> LDTM, TMA loads -> UTMALDG, TMA stores -> UTMASTG, "
                 "tcgen05.commit -> UTCBAR; HMMA / HGMMA (legacy mma.sync / Hopper wgmma) must be 0.")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[-14:]))


if __name__ == "__main__":
    main()
