#!/usr/bin/env python
"""Aggregate an ncu report's per-SASS-instruction samples / instruction counts by CUDA source line.

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep <kernel regex> <mangled-name substring> [top] [inst]

ncu's CSV source page carries no line numbers, so the i-th SASS instruction of the kernel is matched
with the i-th instruction of the same function in `nvdisasm -g` of the cubin inside libpfb.so
(the library must be the build that was profiled)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(mangled_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "pyrecest_b200", "lib", "libpfb.so")], cwd=tmp,
                   capture_output=True)
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur, line, out, active = None, None, [], False
        for ln in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                if active and out:
                    return out
                active = mangled_sub in m.group(1)
                out = []
                continue
            if not active:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur, line = os.path.basename(m.group(1)), int(m.group(2))
                continue
            if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
                out.append((cur, line, ln.split("*/", 1)[1].strip()))
        if active and out:
            return out
    raise SystemExit("function not found in any cubin: " + mangled_sub)


def main():
    rep, kre, sub = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    by = 1 if (len(sys.argv) > 5 and sys.argv[5] == "inst") else 0  # sort by instructions instead of samples
    if rep.endswith(".csv"):  # a source page exported on the GPU box (ncu -i rep --page source --csv): reports > 64 MB do not travel
        csvtxt = open(rep).read()
    else:
        csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                                capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(csvtxt)))
    # several launches may be concatenated; take the first block
    hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    start = hdr_idx[0]
    end = hdr_idx[1] - 1 if len(hdr_idx) > 1 else len(rows)
    hdr = rows[start]
    body = [r for r in rows[start + 1:end] if len(r) == len(hdr)]
    lines = sass_lines(sub)
    if len(lines) != len(body):
        print(f"warning: {len(body)} SASS rows in the report vs {len(lines)} in the cubin (different build?)")
    i_s, i_n, i_src = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") or h.startswith("Stall")]
    agg = collections.defaultdict(lambda: [0, 0])
    tot_s = tot_n = 0
    for k, r in enumerate(body):
        s = int(r[i_s] or 0)
        n = int(r[i_n] or 0)
        key = lines[k][:2] if k < len(lines) else ("?", 0)
        agg[key][0] += s
        agg[key][1] += n
        tot_s += s
        tot_n += n
    print(f"total samples {tot_s}, warp instructions {tot_n}")
    src_cache = {}
    for (f, ln), (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][by])[:top]:
        text = ""
        for d in ("pyrecest_b200/csrc", "include"):
            pth = os.path.join(ROOT, d, f or "")
            if os.path.exists(pth):
                src_cache.setdefault(pth, open(pth).read().splitlines())
                if ln and ln <= len(src_cache[pth]):
                    text = src_cache[pth][ln - 1].strip()[:90]
        print(f"{100 * s / max(tot_s, 1):5.1f}% samples  {100 * n / max(tot_n, 1):5.1f}% inst  {f}:{ln}  {text}")


if __name__ == "__main__":
    main()
