#!/usr/bin/env python
"""Instruction budget of one kernel of an ncu report, grouped by source FILE:LINE ranges given on the command line.

  python tools/ncu_regions.py rep.ncu-rep <kernel regex> <mangled substring> <units> name=file:lo-hi [...]

`units` divides the totals (e.g. the number of particles) so that the output reads "instructions per particle".
Lines that fall in no range are listed individually (top 15)."""
import collections
import csv
import io
import subprocess
import sys

sys.path.insert(0, __import__("os").path.dirname(__file__))
import ncu_lines as N  # noqa: E402


def main():
    rep, kre, sub, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    ranges = []
    for a in sys.argv[5:]:
        name, spec = a.split("=")
        f, lh = spec.split(":")
        lo, hi = lh.split("-")
        ranges.append((name, f, int(lo), int(hi)))
    if rep.endswith(".csv"):
        txt = open(rep).read()
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                             capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hidx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    hdr = rows[hidx[0]]
    end = hidx[1] - 1 if len(hidx) > 1 else len(rows)
    body = [r for r in rows[hidx[0] + 1:end] if len(r) == len(hdr)]
    lines = N.sass_lines(sub)
    i_n, i_t, i_s = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for k, r in enumerate(body):
        f, l, _ = lines[k] if k < len(lines) else ("?", 0, "")
        key = f"{f}:{l}"
        for name, rf, lo, hi in ranges:
            if f == rf and lo <= (l or 0) <= hi:
                key = name
                break
        agg[key][0] += int(r[i_n] or 0)
        agg[key][1] += int(r[i_t] or 0)
        agg[key][2] += int(r[i_s] or 0)
    tot = sum(v[0] for v in agg.values())
    tots = sum(v[2] for v in agg.values())
    print(f"warp instructions {tot}  ({tot * 32 / units:.1f} per unit at 32 lanes; thread instructions {sum(v[1] for v in agg.values()) / units:.1f} per unit)")
    named = {n for n, *_ in ranges}
    shown = 0
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if k not in named:
            shown += 1
            if shown > 15:
                continue
        print(f"{k:42s} {100 * v[0] / tot:5.1f}% inst  {v[0] * 32 / units:7.1f}/unit   {100 * v[2] / max(tots, 1):5.1f}% samples")


if __name__ == "__main__":
    main()
