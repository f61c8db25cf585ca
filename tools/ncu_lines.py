#!/usr/bin/env python
"""Per-source-line executed-instruction and stall-sample totals for one kernel of an .ncu-rep.

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep [kernel-substring] [--so berserkerquake2_b200/libbq2gpu.so]

ncu's CSV source page is per SASS instruction; nvdisasm -g gives the SASS -> file:line map of the same cubin
(the library must be the build that was profiled, compiled with -lineinfo).  Joined by instruction offset.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_lines(so, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    out = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur_fn, line, inl = None, None, None
        for l in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
            if m:
                cur_fn = m.group(1); out.setdefault(cur_fn, {}); continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
            if m:
                line = (os.path.basename(m.group(1)), int(m.group(2))); continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if m and cur_fn:
                out[cur_fn][int(m.group(1), 16)] = (line, m.group(2).strip())
    cands = [(fn, d) for fn, d in out.items() if kernel_sub in fn and d]
    if len(cands) > 1:
        print("# several kernels match %r, taking the first: %s" % (kernel_sub, ", ".join(fn for fn, _ in cands)), file=sys.stderr)
    if cands:
        return cands[0][1]
    raise SystemExit(f"no kernel matching {kernel_sub!r} in {so}")


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "kernel"
    so = sys.argv[sys.argv.index("--so") + 1] if "--so" in sys.argv else "berserkerquake2_b200/libbq2gpu.so"
    # --sass <mangled substring>: which SASS function to map (template instantiations differ only in the mangled name)
    sass_sub = sys.argv[sys.argv.index("--sass") + 1] if "--sass" in sys.argv else sub
    m = sass_lines(so, sass_sub)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    # first kernel block only
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}; blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    blk = next(b for b in blocks if sub in b["name"])
    hdr = blk["rows"][0]
    ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = blk["rows"][1:]
    base = int(data[0][ia], 16)
    per = defaultdict(lambda: [0, 0, defaultdict(int)])
    tot_i = tot_s = 0
    for r in data:
        off = int(r[ia], 16) - base
        line = m.get(off, (None, ""))[0] or ("?", 0)
        n, s = int(r[ie]), int(r[isamp])
        per[line][0] += n; per[line][1] += s
        for c in stall_cols:
            v = int(r[c] or 0)
            if v:
                per[line][2][hdr[c][6:]] += v
        tot_i += n; tot_s += s
    print(f"kernel {blk['name']}: {tot_i} warp instructions, {tot_s} stall samples")
    src = {}
    for (f, ln) in per:
        p = os.path.join("berserkerquake2_b200/csrc", f)
        if f not in src and os.path.exists(p):
            src[f] = open(p).read().splitlines()
    for (f, ln), (n, s, st) in sorted(per.items(), key=lambda kv: -kv[1][0]):
        if not os.environ.get("NCU_LINES_ALL") and n < tot_i * 0.003 and s < tot_s * 0.005:
            continue
        top = ",".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        text = src.get(f, [""] * (ln + 1))[ln - 1].strip()[:70] if ln else ""
        print(f"{100 * n / tot_i:5.1f}% inst {100 * s / max(tot_s, 1):5.1f}% samp  {f}:{ln:<4d} {text:70s} {top}")


if __name__ == "__main__":
    main()
