#!/usr/bin/env python
"""bench.py -- headline benchmark of the BDPT -> histogram -> IR hot path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): procedural furnished room (~100 k triangles, 8 materials,
8 absorption bands), 1 source / 1 listener, 2^20 BDPT path pairs per IR update, max depth 16,
1000 x 1 ms bins, 1 s 48 kHz IR.  A step = one IR update = fs_trace + fs_build_ir (histogram ->
IR -> convolver partition spectra).  N GPUs: weak scaling, every rank traces 2^20 pairs of a
global N * 2^20 work range against a replicated BVH, one NCCL integer reduce per step.

One JSON line on stdout (rank 0).  `value` = path pairs/s with everything resident in HBM;
`e2e` = the same through the C-ABI with host buffers (positions in, histogram + IR out);
`roofline` = k_trace_closest_q (BVH traversal) algorithmic bytes / its CUDA-event time vs the measured HBM
peak; `cpu_baseline` = the CPU oracle (a port of the reference's loop) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "audio-pathtracer_b200"))

import numpy as np  # noqa: E402

METRIC = "bdpt_path_pairs_per_second"
UNIT = "paths/s"
WORKLOAD = "furnished_room_100k_tris_1M_paths_depth16_8bands"
SEED0 = 1000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--paths", type=int, default=1 << 20)
    ap.add_argument("--depth", type=int, default=16)
    ap.add_argument("--scene", default="furnished_room", choices=["furnished_room", "mine_tunnels", "concert_hall", "shoebox"],
                    help="default furnished_room = BASELINE.json configs[1]; the others are extra measurements")
    ap.add_argument("--sources", type=int, default=1,
                    help="emitters (BASELINE configs[3]: 64 in the mine tunnels); --paths stays the path pairs per GPU per step, "
                         "split evenly over the sources")
    ap.add_argument("--share-listener", action="store_true",
                    help="FS_FLAG_SHARE_LISTENER: listener subpaths keyed by the path index, traced once per update and shared "
                         "by all sources (SURVEY 8f rank 4; different random numbers than the default mode)")
    ap.add_argument("--cpu-sample-paths", type=int, default=1 << 19)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def config_dict(args, n):
    wl = WORKLOAD if (args.scene == "furnished_room" and args.paths == 1 << 20 and args.depth == 16 and args.sources == 1) else \
        "%s_%d_paths_depth%d_8bands%s" % (args.scene, args.paths, args.depth, "_%dsources" % args.sources if args.sources > 1 else "")
    return {"workload": wl, "scene": "%s (seeded procedural)" % args.scene, "paths_per_gpu_per_step": args.paths,
            "max_depth": args.depth, "bands": 8, "bins": 1000, "sources": args.sources, "rr_prob": 0.9,
            "parallelism": "path-range sharding x%d, replicated BVH, one int64 reduce" % n,
            "share_listener": bool(args.share_listener),
            "l2": "no explicit flush: per-step wavefront state+records (~0.6 GB) exceed the 126 MB L2; "
                  "the ~11 MB BVH is L2-resident by design, as in steady-state 60 Hz refresh"}


# ---------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:                                   # one long-lived nvidia-smi in loop mode (-lms), parsed at the end
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.3)                    # first sample lands before the timed region starts
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is None:
            return
        time.sleep(0.1)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        for line in out.strip().splitlines():
            self.rows.append([c.strip() for c in line.split(",")])

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def traffic_from_profiles(args):
    """dram__bytes_read.sum + dram__bytes_write.sum per k_trace_closest launch from the committed ncu --set full
    capture of this workload (profiles/*_traffic.json); None for workloads that were not captured"""
    if not (args.scene == "furnished_room" and args.paths == 1 << 20 and args.depth == 16):
        return None
    best = None
    pdir = os.path.join(ROOT, "profiles")
    for f in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if f.endswith("_traffic.json"):
            try:
                best = float(json.load(open(os.path.join(pdir, f)))["dram_bytes_per_launch"])
            except Exception:
                pass
    return best


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md, MEASURED_PEAKS.json absent)"


# ---------------------------------------------------------------------------------------------
# reference arm: the CPU oracle (port of the reference's BDPT loop) on all host threads
# ---------------------------------------------------------------------------------------------
def cpu_oracle_rate(args, n_paths, threads, repeats=1, seed=SEED0):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    from frequensee import scenes
    sc = scenes.by_name(args.scene)
    sc.sources = sc.sources[:args.sources]
    S = po.Scene(sc.verts, sc.tri_mat, sc.absorption, use_bvh=True)
    cfg = po.default_config(flags=po.FLAG_SHARE_LISTENER) if args.share_listener else po.default_config()
    times = []
    for r in range(repeats):
        t0 = time.perf_counter()
        per_source = max(1, n_paths // len(sc.sources))      # n_paths = path pairs over all sources
        h, st = S.trace(cfg, sc.sources, sc.listener, per_source, args.depth, seed + r, n_threads=threads)
        for si in range(len(sc.sources)):
            po.build_ir(cfg, h[si], per_source)
        times.append(time.perf_counter() - t0)
    return times, st


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    sample = min(args.paths, 1 << 17)
    cpu_oracle_rate(args, sample, threads, repeats=max(1, min(args.warmup, 1)))
    times, st = cpu_oracle_rate(args, sample, threads, repeats=args.steps)
    total = sum(times)
    v = sample * args.steps / total
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 + u64 Q32.32", "data": "synthetic",
            "config": dict(config_dict(args, args.gpus), sample_paths_per_step=sample),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d path pairs per step of the same workload (oracle/fs_oracle.c: the reference's "
                                       "BDPT loop restated in C; the UE plugin itself cannot be built)" % sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "mrays_per_s": (st["ext_rays"] + st["shadow_rays"]) / (total / args.steps) / 1e6}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import frequensee as fs
    from frequensee import scenes, capi
    from frequensee.distributed import shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: anything native code prints there (e.g. "NCCL version ...") goes to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, K, W = world, args.steps, args.warmup
    P, D = args.paths, args.depth
    sc = scenes.by_name(args.scene)
    sc.sources = sc.sources[:args.sources]
    NS_ = len(sc.sources)
    if P * N % NS_:
        raise SystemExit("--paths x --gpus must be a multiple of --sources")
    base_flags = capi.FLAG_SHARE_LISTENER if args.share_listener else 0
    ctx = fs.Context(device=local, flags=base_flags)        # production configuration: no per-kernel events, batch lanes on
    ctx.set_scene(sc.verts, sc.tri_mat, sc.absorption)
    stream = torch.cuda.Stream()                            # a real (non-NULL) stream: the C-ABI treats NULL as "own stream"
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    B, Kb = ctx.cfg.n_bands, ctx.cfg.n_bins
    d_hist = torch.zeros((NS_, B, Kb), dtype=torch.int64, device="cuda")
    n_global = P * N // NS_                                 # per-source path count of the whole job
    g_first, g_count = shard_range(n_global * NS_, rank, N) # contiguous range of the global index g = source * n + i

    def step_device(seed):
        ctx.trace_range_device(sc.sources, sc.listener, n_global, g_first, g_count, D, seed, d_hist.data_ptr(), True)
        if N > 1:
            dist.reduce(d_hist, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            ctx.set_histogram_device(d_hist.data_ptr(), NS_, n_global)
            if NS_ == 1:
                ctx.build_ir(0, want_ir=False)
            else:
                ctx.build_ir_all(NS_, want_ir=False)

    def barrier():
        if N > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(W):
        step_device(SEED0 - 1 - w)
    barrier()
    launches0 = ctx.stats()["kernel_launches"]
    ext_ms, con_ms, evl_ms, ext_launches = 0.0, 0.0, 0.0, 0
    rays = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with Clocks(local) as clk:
        barrier()
        e0.record(stream)
        for k in range(K):
            step_device(SEED0 + k)
        e1.record(stream)
        barrier()
        dev_ms = e0.elapsed_time(e1)
    st = ctx.stats()                                        # kernel-class times of the LAST timed step
    launches = st["kernel_launches"] - launches0
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if N > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = P * N * K / (dev_ms * 1e-3)

    # per-kernel-class device time: the same steps (same seeds) re-run on a second context with FS_FLAG_TIME_KERNELS,
    # i.e. CUDA events on the launching stream around every stage and every k_trace_closest launch.  That context runs
    # its batches on ONE lane, so each kernel is timed alone (in the headline run two batch lanes overlap their kernels).
    tctx = fs.Context(device=local, flags=capi.FLAG_TIME_KERNELS | base_flags)
    tctx.set_scene(sc.verts, sc.tri_mat, sc.absorption)
    tctx.set_stream(stream.cuda_stream)
    per = []
    for k in range(-1, min(K, 5)):
        tctx.trace_range_device(sc.sources, sc.listener, n_global, g_first, g_count, D, SEED0 + k, d_hist.data_ptr(), True)
        torch.cuda.synchronize()
        if k >= 0:
            per.append(tctx.stats())
    tctx.close()
    ext_ms = float(np.mean([s["extend_ms"] for s in per])); con_ms = float(np.mean([s["connect_ms"] for s in per]))
    trc_ms = float(np.mean([s["trace_ms"] for s in per]))
    evl_ms = float(np.mean([s["eval_ms"] for s in per])); ext_launches = per[0]["extend_launches"]
    rays = float(np.mean([s["ext_rays"] + s["shadow_rays"] for s in per]))
    ext_rays = float(np.mean([s["ext_rays"] for s in per]))
    connected = float(np.mean([s["connected"] for s in per]))

    # algorithmic bytes: visit counts of the same rays from the instrumented build of the same kernels
    cctx = fs.Context(device=local, flags=capi.FLAG_COUNT_VISITS | base_flags)
    cctx.set_scene(sc.verts, sc.tri_mat, sc.absorption)
    cnt = []
    for k in range(min(K, 2)):
        cctx.trace_range(sc.sources, sc.listener, n_global, g_first, g_count, D, SEED0 + k)
        cnt.append(cctx.stats())
    cctx.close()
    en = float(np.mean([c["node_visits"] - c["shadow_node_visits"] for c in cnt]))
    et = float(np.mean([c["tri_tests"] - c["shadow_tri_tests"] for c in cnt]))
    er = float(np.mean([c["ext_rays"] for c in cnt]))
    # SURVEY.md 8(d) per-ray figure with this build's record sizes:
    # bytes_ray = 64 * n_node (BVH2 node, both child boxes) + 48 * n_tri (v0,e1,e2) + 32 (ray record read) + 8 (hit write)
    ext_bytes = 64.0 * en + 48.0 * et + 40.0 * er
    peak, peak_src = measured_peak()
    achieved = ext_bytes / (trc_ms * 1e-3) / 1e9 if trc_ms > 0 else None
    ext_ms_all = ext_ms
    ext_ms = trc_ms
    bvh_mb = (sc.n_tris * 0.5 * 64 + sc.n_tris * 64) / 1e6          # reachable 4-wide nodes (~T/2 x 64 B) + triangles (64 B)
    roofline = {"bound": "hbm", "kernel": "k_trace_closest_q (persistent 4-wide BVH closest-hit traversal, warp-shared triangle "
                                          "queue), %d launches per step on the one-lane timing context" % ext_launches,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "traffic": traffic_from_profiles(args), "peak_source": peak_src,
                "bytes_per_launch": ext_bytes / max(ext_launches, 1), "ms_per_launch": ext_ms / max(ext_launches, 1),
                "nodes_per_ray": en / er, "tris_per_ray": et / er,
                "note": ("nodes + triangles (~%.0f MB) are L2-resident: achieved is algorithmic fetch bandwidth, served mostly "
                         "by L2/L1, normalised by the measured HBM copy peak as SURVEY.md 8(d) prescribes" % bvh_mb) if bvh_mb < 100
                        else ("nodes + triangles (~%.0f MB) exceed the 126 MB L2: algorithmic fetch bandwidth against the "
                              "measured HBM copy peak; n_node / n_tri counted by the instrumented build of the same kernel" % bvh_mb)}

    # e2e: the public C-ABI call with host buffers in and out, copies inside the timed region
    ctx.set_stream(None)
    e2e = None
    if rank == 0 or N > 1:
        def step_host(seed):
            if N == 1:
                h = ctx.trace(sc.sources, sc.listener, n_global, D, seed)    # positions H2D, histogram D2H
                ir = ctx.build_ir(0) if NS_ == 1 else ctx.build_ir_all(NS_)  # IR D2H
                return h, ir
            h = ctx.trace_range(sc.sources, sc.listener, n_global, g_first, g_count, D, seed)
            return h, None
        step_host(SEED0 - 1)
        barrier()
        t0 = time.perf_counter()
        for k in range(K):
            h, ir = step_host(SEED0 + k)
            if N > 1:                                                         # host-buffer API: reduce through the device tensor
                d_hist.copy_(torch.from_numpy(h.view(np.int64)))
                dist.reduce(d_hist, dst=0, op=dist.ReduceOp.SUM)
                if rank == 0:
                    ctx.set_histogram(d_hist.cpu().numpy().view(np.uint64), n_global)
                    ir = ctx.build_ir(0) if NS_ == 1 else ctx.build_ir_all(NS_)
        barrier()
        wall = time.perf_counter() - t0
        tw = torch.tensor([wall], dtype=torch.float64, device="cuda")
        if N > 1:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        wall = float(tw.item())
        hb = NS_ * B * Kb * 8
        e2e = {"value": P * N * K / wall, "unit": UNIT, "h2d_bytes_per_step": int(sc.sources.nbytes + sc.listener.nbytes),
               "d2h_bytes_per_step": int(hb + NS_ * ctx.cfg.n_channels * ctx.cfg.sample_rate * 4), "ms_per_step": 1e3 * wall / K}
    cpu = None
    if rank == 0 and N == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        times, so = cpu_oracle_rate(args, args.cpu_sample_paths, threads, repeats=1)
        cpu = {"value": args.cpu_sample_paths / times[0], "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d path pairs of the same workload (oracle/fs_oracle.c, pthreads, %d threads), %.2f s"
                         % (args.cpu_sample_paths, threads, times[0])}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": N, "steps": K, "warmup": W,
                "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 + u64 Q32.32", "data": "synthetic", "config": config_dict(args, N),
                "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "cpu_baseline": cpu,
                "mrays_per_s": rays * N / (dev_ms / K * 1e-3) / 1e6, "ms_per_ir_update": dev_ms / K,
                "stage_ms": {"trace_closest": trc_ms, "shade_gen": ext_ms_all - trc_ms, "connect": con_ms, "eval_splat": evl_ms},
                "rays_per_step_per_gpu": rays, "ext_rays_per_step_per_gpu": ext_rays, "connected_per_step_per_gpu": connected,
                "triangles": int(sc.n_tris)}
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    ctx.close()
    if N > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
