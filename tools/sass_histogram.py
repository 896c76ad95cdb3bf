"""Per-kernel SASS opcode evidence of libteam_b200.so (runs without a GPU):
    python tools/sass_histogram.py > profiles/r2_sass_histogram.md
Counts, for every kernel in the library, the Blackwell-specific mnemonics (/opt/skills/guides/B200_PROFILING.md):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA, LDGMC/multimem, HMMA = mma.sync, SYNCS = mbarrier,
plus instruction count."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "team-temporal-evolution-aware-multimodal-model_b200", "libteam_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
KEYS = (("UTC*MMA (tcgen05.mma)", r"\bUTC\w*MMA"), ("LDTM/STTM (tcgen05.ld/st)", r"\b(LDTM|STTM)"), ("UTMALDG/UTMASTG/UBLKCP (TMA)", r"\b(UTMA(LDG|STG)|UBLKCP)"),
        ("UTCBAR / SYNCS (mbarrier)", r"\b(UTCBAR|SYNCS)"), ("multimem / LDGMC", r"\b(LDGMC|REDGMC|STGMC|multimem)"), ("HMMA (mma.sync)", r"\bHMMA"),
        ("SHFL", r"\bSHFL"), ("LDG.128 / STG.128", r"\b(LDG|STG)\.E\.(\w+\.)*128"))
cur, rows = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        rows[cur] = collections.Counter()
        continue
    if cur is None or "/*" not in line:
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(.*?);", line)
    if not m:
        continue
    ins = m.group(1)
    rows[cur]["instructions"] += 1
    for k, pat in KEYS:
        if re.search(pat, ins):
            rows[cur][k] += 1
print("# SASS opcode histogram per kernel (`cuobjdump -sass libteam_b200.so`, sm_100a)\n")
print("| kernel | instr | " + " | ".join(k for k, _ in KEYS) + " |")
print("|---|---|" + "---|" * len(KEYS))
tot = collections.Counter()
for fn, c in rows.items():
    name = demangle(fn)
    name = re.sub(r"\(.*", "", name).replace("team::", "").replace("void ", "")
    print(f"| `{name}` | {c['instructions']} | " + " | ".join(str(c[k]) if c[k] else "" for k, _ in KEYS) + " |")
    tot.update(c)
print(f"| **total ({len(rows)} kernels)** | {tot['instructions']} | " + " | ".join(str(tot[k]) for k, _ in KEYS) + " |")
