#!/usr/bin/env python
"""What the compiled library contains, as evidence for profiles/: per kernel of libbq2gpu.so the SASS instruction count and
the counts of the mnemonics that identify the B200 paths used (UBLKCP = cp.async.bulk / TMA bulk copy, SYNCS = mbarrier,
ACQBULK = programmatic-dependent-launch wait, VOTE / REDUX / POPC = the warp-level compaction, MUFU = the rsqrt / rcp of
the exact ratio sequence), registers / shared memory / spills from ptxas, and the target architecture of every cubin.

  python tools/sass_summary.py > profiles/r02_sass_summary.md"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "berserkerquake2_b200", "libbq2gpu.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
archs = Counter(re.findall(r"arch = (sm_\w+)", txt))
kern, cur = OrderedDict(), None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); kern[cur] = Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        kern[cur]["_n"] += 1
        kern[cur][op.split(".")[0]] += 1
        if op.startswith("SYNCS") or op.startswith("MUFU") or op.startswith("STG") or op.startswith("LDG") or op.startswith("LDS") or op.startswith("STS"):
            kern[cur][op] += 1
dem = subprocess.run(["c++filt"] + list(kern), capture_output=True, text=True).stdout.splitlines()
log = os.path.join(ROOT, "berserkerquake2_b200", "csrc", "build", "bq2_kernels.ptxas.log")
res = {}
if os.path.exists(log):
    L = open(log).read().splitlines()
    for i, l in enumerate(L):
        m = re.search(r"Compiling entry function '(\S+)' for '(sm_\w+)'", l)
        if m:
            blob = " ".join(L[i:i + 4])
            r = re.search(r"Used (\d+) registers", blob); sp = re.search(r"(\d+) bytes spill stores", blob); sm = re.search(r"(\d+) bytes smem", blob)
            res[m.group(1)] = (r.group(1) if r else "?", sp.group(1) if sp else "?", sm.group(1) if sm else "0", m.group(2))
print("# Round 2: SASS of `berserkerquake2_b200/libbq2gpu.so` (`tools/sass_summary.py`: `cuobjdump -sass` + the ptxas log)")
print()
print("Cubins in the library: " + ", ".join(f"{a} x{n}" for a, n in archs.items()) + " -- nothing else (no PTX-only or other-arch code).")
print("`UBLKCP` = `cp.async.bulk.shared::cluster.global` (TMA bulk copy), `SYNCS.*` = mbarrier operations, `ACQBULK` = `griddepcontrol.wait`")
print("(programmatic dependent launch), `VOTE` / `POPC` / `REDUX` = ballot compaction and warp reductions, `MUFU.RSQ` / `MUFU.RCP` = the")
print("straight-line `num / sqrtf(len2)` sequence.  No `HMMA` / `UTCMMA` / `TCGEN05`: nothing on this path is a contraction.")
print()
print("| kernel | SASS instr. | regs | spill B | static smem B | UBLKCP | SYNCS | ACQBULK | VOTE | POPC | REDUX | MUFU | LDS | STS | LDG | STG | tensor ops |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for (mang, c), name in zip(kern.items(), dem):
    short = re.sub(r"\(anonymous namespace\)::", "", name)
    short = re.sub(r"\(.*$", "", short).replace("void ", "")
    r = res.get(mang, ("?", "?", "?", "?"))
    tens = sum(v for k, v in c.items() if k.startswith(("HMMA", "UTCMMA", "TCGEN", "IMMA", "QMMA", "UTCHMMA")))
    print(f"| `{short}` | {c['_n']} | {r[0]} | {r[1]} | {r[2]} | {c['UBLKCP']} | {c['SYNCS']} | {c['ACQBULK']} | {c['VOTE'] + c['VOTEU']} | {c['POPC'] + c['UPOPC']} | "
          f"{c['REDUX']} | {c['MUFU']} | {c['LDS']} | {c['STS']} | {c['LDG']} | {c['STG']} | {tens} |")
print()
w = [c for (m, c) in kern.items() if "bq2_frame_warp_kernelILb0ELb0E" in m]
if w:
    c = w[0]
    det = ", ".join(f"`{k}` x{v}" for k, v in sorted(c.items()) if k.startswith(("SYNCS.", "MUFU.", "STG.", "LDG.", "LDS.", "STS.")))
    print("`bq2_frame_warp_kernel<0,0>` (the kernel of the headline number) in detail: " + det + ".")
