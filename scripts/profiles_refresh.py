#!/usr/bin/env python
"""Turn the outputs of scripts/ncu_passes.sh + a bench.py run (gpurun_out/) into the committed files under profiles/.
    python scripts/profiles_refresh.py <bench-profile-name.json>"""
import csv
import json
import re
import subprocess
import sys

out_name = sys.argv[1] if len(sys.argv) > 1 else "r2_bench_n1_latest.json"
subprocess.run([sys.executable, "scripts/ncu_summarise.py", "gpurun_out/step_metrics.csv", "profiles/r2_ncu_step_summary.txt",
                "profiles/r2_ncu_gemm_traffic.json"], check=True)
rows = list(csv.reader(open("gpurun_out/prof_calls_raw.csv")))
hdr = rows[0]
ki = hdr.index("Kernel Name")
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]
idx = [hdr.index(w) for w in want if w in hdr]
out = [[hdr[ki]] + [hdr[i] for i in idx], [""] + [rows[1][i] for i in idx]]
tot = 0.0
for r in rows[2:]:
    out.append([r[ki][:70]] + [r[i] for i in idx])
    tot += float(r[idx[0]])
csv.writer(open("profiles/r2_ncu_full_calls_c320_c640.csv", "w")).writerows(out)
for r in out[2:]:
    print(r[0][:50].ljust(50), r[1][:8], "us")
print("sum", round(tot, 1), "us over", len(out) - 2, "kernels")
rows = list(csv.reader(line for line in open("gpurun_out/launches.csv") if not line.startswith("==")))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
o = []
for r in rows[1:]:
    if len(r) < len(hdr) or r[col["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("nmm::", "")[:70]
    o.append((r[col["ID"]], name, r[col["Grid Size"]].replace(",", " "), r[col["Block Size"]].replace(",", " "), r[col["Metric Unit"]], r[col["Metric Value"]]))
with open("profiles/r2_ncu_launch_list_bench.csv", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'linear_tc|linear_simt|fused_module|gn_|layernorm_pe|temporal_attention|cfg_ddim' "
            "-c 600 --csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-clips --no-eager\n")
    f.write("# first 600 launches of this library in the bench command (warm-up steps, then the timed step); cold-cache / serialised, not bench values\n")
    f.write("id,kernel,grid,block,unit,duration\n")
    for x in o:
        f.write(",".join('"%s"' % v if "," in v else v for v in x) + "\n")
print(len(o), "launches listed")
d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
open("profiles/" + out_name, "w").write(json.dumps(d) + "\n")
print("bench:", round(d["value"], 1), "TFLOP/s", round(d["ms_per_step"], 3), "ms/step; e2e", round(d["e2e"]["value"], 1), "; GEMM", round(d["roofline"]["achieved"], 1),
      "frac", round(d["roofline"]["frac"], 3), "; clips/s", round(d["clips"]["value"], 2) if d.get("clips") else None, d["clocks"], "launches", d["gpu_launches"])
