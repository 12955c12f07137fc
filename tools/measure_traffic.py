#!/usr/bin/env python
"""Regenerate profiles/traffic_r02.json and profiles/eval_pipe_r02.json: DRAM bytes of one bench step and the FP64
pipe figures of the evaluation kernel, from ncu captures of the bench command itself (run on the GPU box):

    python tools/measure_traffic.py            # classic pipeline (default)

bench.py reads the two files and says so in `roofline.traffic_source`; it never measures under a profiler."""
import csv, datetime, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
rep = os.path.join(OUT, "traffic_capture")
cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed_pipe_fp64.sum,"
       "smsp__thread_inst_executed_pipe_fp64_pred_on.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
       "--clock-control", "none", "-s", "12", "-c", "4", "--csv", "--log-file", rep + ".csv",
       sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3", "--no-cpu-baseline", "--no-side-configs", "--e2e-steps", "1"]
subprocess.run(cmd, check=False, cwd=ROOT, stdout=subprocess.DEVNULL)
rows = [r for r in csv.reader(open(rep + ".csv")) if len(r) > 10]
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
idi = hdr.index("ID")
per = {}
for r in rows[1:]:
    per.setdefault((r[idi], r[ki]), {})[r[mi]] = float(r[vi].replace(",", ""))
# one step = the launches between two evaluation kernels
ids = sorted(per, key=lambda k: int(k[0]))
start = next(i for i, k in enumerate(ids) if "eval_kernel" in k[1])
step = [ids[start]]
for k in ids[start + 1:]:
    if "eval_kernel" in k[1]:
        break
    step.append(k)
tot = sum(per[k]["dram__bytes_read.sum"] + per[k]["dram__bytes_write.sum"] for k in step)
ev = per[step[0]]
n_el = 1_000_000
when = datetime.datetime.utcnow().strftime("%Y-%m-%d")
json.dump({"elements": n_el, "pipeline": "classic", "dram_bytes_per_step": tot,
           "kernels": [{"name": k[1].split("(")[0], "dram_read": per[k]["dram__bytes_read.sum"], "dram_write": per[k]["dram__bytes_write.sum"],
                        "duration_ns_under_ncu": per[k]["gpu__time_duration.sum"]} for k in step],
           "source": f"profiles/traffic_r02.json: tools/measure_traffic.py, {when}"},
          open(os.path.join(ROOT, "profiles", "traffic_r02.json"), "w"), indent=1)
json.dump({"fp64_pipe_busy_pct": ev.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
           "fp64_thread_inst_per_element": ev.get("smsp__thread_inst_executed_pipe_fp64_pred_on.sum", 0.0) / n_el,
           "fp64_warp_inst": ev.get("smsp__inst_executed_pipe_fp64.sum"), "kernel": step[0][1].split("(")[0],
           "source": f"tools/measure_traffic.py, {when}"},
          open(os.path.join(ROOT, "profiles", "eval_pipe_r02.json"), "w"), indent=1)
import shutil
for name in ("traffic_r02.json", "eval_pipe_r02.json"):      # gpurun merges gpurun_out/ back, not profiles/
    shutil.copy(os.path.join(ROOT, "profiles", name), os.path.join(OUT, name))
print(open(os.path.join(ROOT, "profiles", "traffic_r02.json")).read())
print(open(os.path.join(ROOT, "profiles", "eval_pipe_r02.json")).read())
