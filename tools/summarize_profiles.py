"""Turns gpurun_out/{launches_*.csv, prof_*.ncu-rep} into the committed summaries under profiles/.
usage: python tools/summarize_profiles.py <launches.csv> <prof.ncu-rep> <tag> [scene=room|hall] [launches_per_step=16]

<tag>_traffic.json (read by bench.py's roofline block) holds, for the closest-hit instance of k_trace_q over the LAST complete
step that was captured: mean DRAM bytes per launch, time-weighted lts__throughput, and the thread-level / warp-level
instruction totals of that step (bench.py divides them by the extension rays it measures)."""
import csv, collections, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launches, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
scene = sys.argv[4] if len(sys.argv) > 4 else "room"
per_step = int(sys.argv[5]) if len(sys.argv) > 5 else 16
out_md = os.path.join(ROOT, "profiles", tag + "_summary.md")
lines = ["# ncu summary `%s`" % tag, "",
         "Command: `python bench.py --steps 2 --warmup 1 --no-cpu-baseline` (furnished room, 2^20 path pairs, depth 16).",
         "Launch list: `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES).", ""]
rows = list(csv.reader(open(launches)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]; kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
seq = []
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = r[kn].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    agg[name][0] += 1; agg[name][1] += v; seq.append((name, v))
# one IR update = from the last k_reset_queues(pair) to k_ir_spectra: take the last complete step
idx = [i for i, (n, _) in enumerate(seq) if n == "k_ir_spectra"]
step = collections.defaultdict(lambda: [0, 0.0])
if len(idx) >= 2:
    for n, v in seq[idx[-2] + 1: idx[-1] + 1]:
        step[n][0] += 1; step[n][1] += v
tot = sum(v[1] for v in step.values()) or 1.0
lines += ["## Kernel shares of one IR update (last complete step of the run)", "", "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
for k, v in sorted(step.items(), key=lambda kv: -kv[1][1]):
    lines.append("| %s | %d | %.1f | %.1f %% |" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot))
lines += ["", "## Whole run, all kernels", "", "| kernel | launches | total us |", "|---|---:|---:|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append("| %s | %d | %.1f |" % (k, v[0], v[1] / 1e3))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]; ix = {n: i for i, n in enumerate(h)}
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio",
        "smsp__warps_eligible.avg.per_cycle_active"]
lines += ["", "## `ncu --set full` of the dominant kernel (k_trace_closest), %d launches captured" % (len(rr) - 2), ""]
lines.append("| metric | " + " | ".join("launch %d" % i for i in range(len(rr) - 2)) + " | unit |")
lines.append("|---|" + "---:|" * (len(rr) - 2) + "---|")
for w in want:
    if w in ix:
        lines.append("| %s | %s | %s |" % (w, " | ".join(r[ix[w]] for r in rr[2:]), rr[1][ix[w]]))
def tobytes(val, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(val.replace(",", "")) * m.get(unit, 1)
def num(r, k):
    return float(r[ix[k]].replace(",", ""))
# the closest-hit instance only (template arguments <COUNT, TEX, ANY = false>), last complete step
closest = [r for r in rr[2:] if "k_trace_q" in r[ix["Kernel Name"]] and not r[ix["Kernel Name"]].rstrip().endswith("true>") and "(bool)1>" not in r[ix["Kernel Name"]].replace(" ", "")[-9:]]
if not closest:
    closest = rr[2:]
step_rows = closest[-per_step:] if len(closest) >= per_step else closest
dur = [num(r, "gpu__time_duration.sum") for r in step_rows]
dram = [tobytes(r[ix["dram__bytes_read.sum"]], rr[1][ix["dram__bytes_read.sum"]]) + tobytes(r[ix["dram__bytes_write.sum"]], rr[1][ix["dram__bytes_write.sum"]]) for r in step_rows]
lts = [num(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed") for r in step_rows]
tinst = [num(r, "smsp__inst_executed.sum") * num(r, "smsp__thread_inst_executed_per_inst_executed.ratio") for r in step_rows]
winst = [num(r, "smsp__inst_executed.sum") for r in step_rows]
json.dump({"kernel": "k_trace_q<closest>", "scene": scene, "launches": len(step_rows), "launches_per_step": per_step,
           "dram_bytes_per_launch": sum(dram) / len(dram),
           "lts_throughput_pct": sum(l * d for l, d in zip(lts, dur)) / sum(dur),
           "thread_inst_per_step": sum(tinst), "warp_inst_per_step": sum(winst),
           "active_lanes_per_inst": sum(tinst) / sum(winst),
           "source": os.path.basename(rep)}, open(os.path.join(ROOT, "profiles", tag + "_traffic.json"), "w"))
open(out_md, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:40]))
