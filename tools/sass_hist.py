"""SASS opcode histogram per kernel of libmdf_b200.so: how many tcgen05 MMAs (UTC*MMA), TMEM loads / stores (LDTM / STTM), TMA
copies (UTMALDG = tensor-map form, UBLKCP = 1-D bulk form), mbarrier operations (SYNCS) and tensor-memory allocations
(UTCATOMSWS / UTCBAR ...) each kernel's machine code holds.  Evidence that the contractions run on the Blackwell tensor pipe.

  python tools/sass_hist.py [--out profiles/r02_sass_histogram.md]
"""
import argparse
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GROUPS = [("tcgen05.mma", re.compile(r"^UTC[A-Z]*MMA")), ("tcgen05.ld", re.compile(r"^LDTM")), ("tcgen05.st", re.compile(r"^STTM")),
          ("TMA tensor", re.compile(r"^UTMA(LDG|STG|PF|REDG)")), ("TMA bulk", re.compile(r"^UBLK(CP|RED|PF)")),
          ("mbarrier", re.compile(r"^SYNCS")), ("tcgen05 ctl", re.compile(r"^UTC(BAR|ATOMSWS|CP|SHIFT)")),
          ("FFMA/FMUL/FADD", re.compile(r"^(FFMA|FMUL|FADD)")), ("HMMA (mma.sync)", re.compile(r"^HMMA"))]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "metagenomic-deepfri_b200", "libmdf_b200.so"))
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_sass_histogram.md"))
    args = ap.parse_args()
    sass = subprocess.run(["cuobjdump", "-sass", args.lib], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur["_total"] += 1
            for name, rx in GROUPS:
                if rx.match(m.group(1)):
                    cur[name] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    rows = []
    for (mangled, c), name in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("mdf::", "")
        rows.append((short, c))
    rows.sort(key=lambda r: (-r[1]["tcgen05.mma"], -r[1]["_total"]))
    head = ["kernel", "SASS instr"] + [g[0] for g in GROUPS]
    out = ["# SASS opcode histogram of libmdf_b200.so (sm_100a), one row per kernel", "",
           "`python tools/sass_hist.py` (cuobjdump -sass).  `tcgen05.mma` = UTC*MMA opcodes, `tcgen05.ld/st` = LDTM / STTM, TMA = UTMALDG (tensor-map) "
           "and UBLKCP (1-D bulk); no kernel uses `mma.sync` (HMMA).", "", "| " + " | ".join(head) + " |", "|" + "---|" * len(head)]
    for short, c in rows:
        out.append("| `" + short + "` | " + " | ".join(str(c[k]) for k in ["_total"] + [g[0] for g in GROUPS]) + " |")
    tot = collections.Counter()
    for _, c in rows:
        tot.update(c)
    out.append("| **all kernels** | " + " | ".join(str(tot[k]) for k in ["_total"] + [g[0] for g in GROUPS]) + " |")
    open(args.out, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:40]))


if __name__ == "__main__":
    main()
