"""Count the SASS mnemonics that prove tcgen05 / TMEM / TMA / mma.sync use, per kernel of the embedded cubin
(B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP).  Runs on the
CPU (cuobjdump only).  usage: python tools/sass_evidence.py [out.md]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUBIN = os.path.join(ROOT, "aule-attention_b200", "build", "aule_kernels.cubin")
PATS = [("UTC*MMA (tcgen05.mma)", r"\bUTC\w*MMA"), ("LDTM (tcgen05.ld)", r"\bLDTM"), ("STTM (tcgen05.st)", r"\bSTTM"),
        ("UTMALDG (TMA load)", r"\bUTMALDG"), ("UTMASTG (TMA store)", r"\bUTMASTG"), ("UBLKCP (bulk copy)", r"\bUBLKCP"),
        ("UTMAPF/prefetch", r"\bUTMAPF|\bUTMACCTL"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("HMMA (mma.sync)", r"\bHMMA"), ("LDSM (ldmatrix)", r"\bLDSM"),
        ("MUFU.EX2", r"MUFU\.EX2"), ("FFMA2/FADD2/FMUL2 (f32x2)", r"\bF(FMA|ADD|MUL)2\b")]
out = subprocess.run(["cuobjdump", "-sass", CUBIN], capture_output=True, text=True, check=True).stdout
counts, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur:
        for name, pat in PATS:
            if re.search(pat, line):
                counts[cur][name] += 1
rows = ["# SASS evidence per kernel (`cuobjdump -sass aule-attention_b200/build/aule_kernels.cubin`, sm_100a)", "",
        "Static instruction counts (not executed counts).  tcgen05.mma appears as `UTC*MMA`, tcgen05.ld/st as `LDTM`/`STTM`, "
        "`cp.async.bulk.tensor` as `UTMALDG`/`UTMASTG`, `cp.async.bulk` as `UBLKCP`, mbarrier ops as `SYNCS`; the paged decode "
        "kernel uses `HMMA` (mma.sync) + `LDSM` (ldmatrix) by design (DESIGN.md 4.4).", "",
        "| kernel | " + " | ".join(n for n, _ in PATS) + " |", "|---|" + "---|" * len(PATS)]
for k, c in counts.items():
    if "sm100" in k or "paged" in k:
        rows.append(f"| `{k}` | " + " | ".join(str(c.get(n, 0)) for n, _ in PATS) + " |")
text = "\n".join(rows) + "\n"
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
print(text)
