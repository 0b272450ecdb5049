"""python tools/sass_summary.py > profiles/r2_sass_summary.md : `cuobjdump -sass` mnemonic counts per kernel of the built library
(which kernels really are tcgen05 / TMA / mma.sync code).  CPU only."""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(REPO, "cpc_audio_b200", "libcpc_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
COLS = [("UTCHMMA", r"\bUTCHMMA"), ("UTMALDG", r"\bUTMALDG"), ("UBLKCP/UBLKRED/UTMAREDG", r"\b(UBLKCP|UBLKRED|UTMAREDG)"),
        ("LDTM", r"\bLDTM"), ("HMMA", r"\bHMMA"), ("SYNCS", r"\bSYNCS"), ("LDGSTS", r"\bLDGSTS"), ("REDG/ATOMG", r"\b(REDG|ATOMG|ATOM)\.")]
rows = collections.OrderedDict()
blocks = re.split(r"\n\s*Function : ", sass)[1:]
for blk, nm in zip(blocks, names):
    nm = re.sub(r"^void ", "", nm)
    nm = re.sub(r"\(anonymous namespace\)::|cpcb200::", "", nm)
    nm = re.sub(r"\(.*$", "", nm).replace("__nv_bfloat16", "bf16")
    instr = len(re.findall(r"/\*[0-9a-f]{4,6}\*/", blk))
    rows[nm] = [instr] + [len(re.findall(rx, blk)) for _, rx in COLS]
print("# SASS summary of libcpc_b200.so (round 2, final kernels): `cuobjdump -sass` mnemonic counts per kernel\n")
print("Built with `nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo`; regenerate with `python tools/sass_summary.py`.  "
      "`UTCHMMA` = tcgen05.mma, `UTMALDG` = TMA tensor load, `LDTM` = tcgen05.ld (TMEM read-out), `UBLKCP` / `UBLKRED` / `UTMAREDG` = "
      "bulk copy / bulk reduce-add, `HMMA` = mma.sync, `SYNCS` = mbarrier, `LDGSTS` = cp.async, `REDG` / `ATOMG` = global reductions / atomics.\n")
print("| kernel | SASS instr | " + " | ".join(c for c, _ in COLS) + " |")
print("|---|---|" + "---|" * len(COLS))
for nm in sorted(rows):
    r = rows[nm]
    print(f"| `{nm}` | {r[0]} | " + " | ".join(str(v) if v else "" for v in r[1:]) + " |")
