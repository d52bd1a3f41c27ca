"""Turns a wide raw CSV of ncu (`ncu --metrics ... --csv --page raw --log-file x.csv`, see tools/ncu_capture.sh) over the
k_trace_q launches of two steps into the committed evidence under profiles/:
  <tag>_<scene>_trace_launches.csv   one row per launch, the metrics only (device attribute columns dropped)
  <tag>_<scene>_traffic.json         read by bench.py's roofline block: for the closest-hit instance over the LAST complete step --
                                     mean DRAM bytes per launch, time-weighted lts__throughput, thread- and warp-level instruction
                                     totals of the step (bench.py divides them by the extension rays it measures), hit rates
usage: python tools/summarize_raw.py <raw.csv> <tag> <room|hall> <launches_per_step> [kernel = k_trace_q | k_path_q]"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw, tag, scene, per_step = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
kern = sys.argv[5] if len(sys.argv) > 5 else "k_trace_q"
suffix = "" if kern == "k_trace_q" else "_" + kern
rows = [r for r in csv.reader(l for l in open(raw) if l.startswith('"'))]
h, units, data = rows[0], rows[1], rows[2:]
ix = {n: i for i, n in enumerate(h)}
keep = ["ID", "Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "launch__registers_per_thread"]
keep = [k for k in keep if k in ix]
out_csv = os.path.join(ROOT, "profiles", "%s_%s%s_trace_launches.csv" % (tag, scene, suffix))
with open(out_csv, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(keep); w.writerow([units[ix[k]] for k in keep])
    for r in data:
        w.writerow([r[ix[k]].replace("<unnamed>::", "") for k in keep])
def num(r, k):
    return float(r[ix[k]].replace(",", ""))
def to_bytes(r, k):
    return num(r, k) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[ix[k]], 1)
def to_us(r, k):
    return num(r, k) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(units[ix[k]], 1)
closest = ([r for r in data if "k_trace_q" in r[ix["Kernel Name"]] and r[ix["Kernel Name"]].replace(" ", "").endswith("(bool)0>")] if kern == "k_trace_q"
           else [r for r in data if kern in r[ix["Kernel Name"]]]) or data
step = closest[-per_step:]
dur = [to_us(r, "gpu__time_duration.sum") for r in step]
dram = [to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum") for r in step]
winst = [num(r, "smsp__inst_executed.sum") for r in step]
tinst = [w * num(r, "smsp__thread_inst_executed_per_inst_executed.ratio") for w, r in zip(winst, step)]
wavg = lambda k: sum(num(r, k) * d for r, d in zip(step, dur)) / sum(dur)      # noqa: E731
j = {"kernel": "k_trace_q<closest>" if kern == "k_trace_q" else kern, "scene": scene, "launches": len(step), "launches_per_step": per_step,
     "us_per_step_under_ncu": sum(dur), "dram_bytes_per_launch": sum(dram) / len(dram),
     "dram_throughput_pct": wavg("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
     "lts_throughput_pct": wavg("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
     "l1_hit_pct": wavg("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": wavg("lts__t_sector_hit_rate.pct"),
     "issue_active_pct": wavg("smsp__issue_active.avg.pct_of_peak_sustained_active"),
     "alu_pipe_pct": wavg("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
     "thread_inst_per_step": sum(tinst), "warp_inst_per_step": sum(winst), "active_lanes_per_inst": sum(tinst) / sum(winst),
     "registers": num(step[0], "launch__registers_per_thread"),
     "source": "ncu --metrics (tools/ncu_capture.sh), %s" % os.path.basename(raw)}
json.dump(j, open(os.path.join(ROOT, "profiles", "%s_%s%s_traffic.json" % (tag, scene, suffix)), "w"), indent=1)
print(json.dumps(j, indent=1))
