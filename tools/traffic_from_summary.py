#!/usr/bin/env python
"""profiles/traffic.json from the ncu summaries (tools/ncu_summary.py): DRAM bytes per stage of one step, summed over
the LAST launch of every kernel of the stage.  bench.py copies them into roofline.traffic.

  python tools/traffic_from_summary.py key=summary.csv [...]     key = workload:particles:f64:resampling:exact|fast
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGES = {"predict_loglik": r"predict_loglik_kernel|predict_kernel|loglik_kernel",
          "scan": r"tile_|scan_",
          "resample": r"msort|record_|resample|systematic"}


def main():
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        out = json.load(open(path))
    except Exception:
        out = {}
    for arg in sys.argv[1:]:
        key, f = arg.split("=")
        last = {}
        for r in list(csv.reader(open(f)))[1:]:
            last[r[0]] = r
        entry = {}
        kernels = {}
        fused = any("predict_loglik_kernel" in nm for nm in last)
        for stage, pat in STAGES.items():
            if stage == "predict_loglik" and fused:
                pat = r"predict_loglik_kernel"  # (a stand-alone predict_kernel in the capture is the prior's sampling)
            tot = 0.0
            names = []
            for name, r in last.items():
                base = name.split("(")[0]
                if re.search(pat, base) and "wrap_kernel" not in base:
                    tot += (float(r[2]) + float(r[3])) * 1e9
                    names.append(base.strip().replace("void ", ""))
            entry[stage] = tot
            kernels[stage] = names
        entry["source"] = (f"{os.path.relpath(f, ROOT)} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, "
                           "summed over the kernels of each stage: " + " | ".join(f"{s}: {', '.join(k)}" for s, k in kernels.items()) + ")")
        out[key] = entry
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps({k: {s: v[s] for s in STAGES} for k, v in out.items()}, indent=1))


if __name__ == "__main__":
    main()
