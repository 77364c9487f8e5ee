#!/usr/bin/env python
"""Turn an ncu metrics CSV of one bench step (scripts/profile_step.py under
   ncu --profile-from-start off --clock-control none --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tmem.sum --log-file X.csv python scripts/profile_step.py)
into the per-kernel summary committed under profiles/ and the GEMM's DRAM traffic per launch that bench.py reports as roofline.traffic.

    python scripts/ncu_summarise.py gpurun_out/step_metrics.csv profiles/r1_ncu_step_summary.txt profiles/r1_ncu_gemm_traffic.json
"""
import csv
import json
import re
import sys
from collections import defaultdict


def main():
    src, out_txt, out_json = sys.argv[1:4]
    rows = list(csv.reader(line for line in open(src) if not line.startswith("==")))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    per = defaultdict(lambda: defaultdict(float))
    launches = defaultdict(set)
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        name = re.sub(r"<.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("nmm::", "").split("(")[0]
        metric, unit, val = r[col["Metric Name"]], r[col["Metric Unit"]], float(r[col["Metric Value"]].replace(",", ""))
        if metric == "gpu__time_duration.sum":
            val *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)          # -> us
        if metric.startswith("dram__bytes"):
            val *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        per[name][metric] += val
        launches[name].add(r[col["ID"]])
    total_us = sum(v["gpu__time_duration.sum"] for v in per.values())
    n_launch = sum(len(v) for v in launches.values())
    lines = [f"one bench step under ncu (cold caches, serialised launches): {total_us / 1e3:.3f} ms over {n_launch} launches",
             "per-launch times are NOT bench values; the kernel's share of the step is what must agree with bench.py"]
    for name, v in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        n = len(launches[name])
        t = v["gpu__time_duration.sum"]
        rd, wr = v.get("dram__bytes_read.sum", 0.0), v.get("dram__bytes_write.sum", 0.0)
        tens = v.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) / max(n, 1)
        lines.append(f"{name:40s} {n:4d} launches {t / 1e3:8.3f} ms ({100 * t / total_us:5.1f}% of step)  dram read {rd / 1e6:9.1f} MB  "
                     f"write {wr / 1e6:9.1f} MB  per launch {(rd + wr) / max(n, 1) / 1e6:7.2f} MB  {(rd + wr) / t / 1e3 if t else 0:7.1f} GB/s  "
                     f"tensor-active {tens:5.1f}%  tmem-inst {v.get('sm__inst_executed_pipe_tmem.sum', 0.0) / max(n, 1):9.0f}/launch")
    open(out_txt, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    g = per.get("linear_tc_kernel")
    if g:
        n = len(launches["linear_tc_kernel"])
        json.dump({"source": f"{out_txt} (ncu metrics pass over one bench step, {n} launches of linear_tc_kernel, cold caches)",
                   "kernel": "linear_tc_kernel", "launches": n,
                   "dram_bytes_read_per_launch": g["dram__bytes_read.sum"] / n, "dram_bytes_write_per_launch": g["dram__bytes_write.sum"] / n,
                   "dram_bytes_per_launch": (g["dram__bytes_read.sum"] + g["dram__bytes_write.sum"]) / n}, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main()
