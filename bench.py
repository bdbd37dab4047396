#!/usr/bin/env python
"""bench.py -- throughput of the path-tracing hot path on N B200s of one node.

Contract (driver): `python bench.py --gpus N --steps K --warmup W` (N>1 under torchrun, one
rank per GPU) prints ONE JSON line from rank 0.

* workload  = BASELINE.json configs[2] -- the configuration the metric is quoted on ("1080p,
  8 bounces"): Disney multi-material scene, 1,310,720 triangles in four BVHTriMesh objects +
  25 analytic spheres + backdrop + floor, three sphere area lights + uniform sky, NEE + MIS,
  PathTracer(8), 1920x1080, 256 spp.  A STEP is one batch of `--spp-per-step` samples per
  pixel per GPU (16 -> 16 steps make the configuration's 256 spp on one GPU).
* value     = Mrays/s: rays actually traced by the GPU kernels (closest-hit path rays incl.
  skip-through, any-hit shadow rays, closest-hit MIS rays; counted by device atomics) summed
  over all ranks / device time of the timed region (CUDA events, max over ranks), scene and
  wavefront state resident in HBM.  spp/s and Mpaths/s ride along.
* e2e       = the same metric through the reference-facing host API,
  CudaPathTracer::Render (ag-pathtracer_b200/host/integrator.h) over HOST accumulator
  buffers: every step copies the float4 accumulator host->device, renders, and copies it
  back device->host inside the timed region.
* roofline  = dominant kernel k_trace_closest: algorithmic bytes (64 B per interior visit +
  48 B per triangle test + 32 B per analytic record + 64 B per ray)
  / its device time, against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
* cpu_baseline = the reference's own CPU integrator (oracle/_ref, compiled from
  /root/reference) on all host cores, on a bounded sample of the same workload.

`--impl reference` times that CPU reference instead (rank 0 only).
Inputs are larger than L2 (scene 215 MB + 1.8 GB of wavefront state vs 126 MB), so no L2
flush is needed between timed iterations.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIG = 3            # BASELINE.json configs[2]
METRIC = "Mrays/s (path+shadow+MIS rays traced, 1080p, 8 bounces)"


def load_pkg():
    import __graft_entry__ as ge
    return ge._load_pkg()


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.device)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for k, n in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


_REF_SCENES = {}


def cpu_reference_run(defaults, level, target_seconds, threads):
    """Time the reference's own CPU integrator (oracle/_ref) on a bounded sample of the
    workload: a centred crop of the 1080p film at 1 spp, grown until it costs ~target_seconds."""
    from oracle import ref_binding as ref
    if not ref.available():
        return None
    W, H = defaults["width"], defaults["height"]
    rs = _REF_SCENES.get(level)
    if rs is None:
        rs = _REF_SCENES[level] = ref.RefScene(CONFIG, level)      # BVH build excluded from the timing, as on the GPU side
        rs.count_rays()
    threads = threads or os.cpu_count() or 1

    def crop_run(cw, ch, spp):
        x0, y0 = (W - cw) // 2, (H - ch) // 2
        ref.ray_counts(reset=True)
        t0 = time.perf_counter()
        _, paths = rs.render(W, H, 0, spp, defaults["max_depth"], defaults["depth_arg"], threads=threads, crop=(x0, y0, x0 + cw, y0 + ch))
        dt = time.perf_counter() - t0
        rc = ref.ray_counts(reset=True)
        return paths, rc["closest"] + rc["any"], dt

    paths, rays, dt = crop_run(320, 180, 1)                       # probe
    rate = paths / max(dt, 1e-6)
    want = max(int(rate * target_seconds), 320 * 180)
    spp = 1
    frac = min(1.0, (want / (W * H)) ** 0.5)
    cw, ch = max(64, int(W * frac) // 16 * 16), max(36, int(H * frac) // 9 * 9)
    if want > W * H:
        spp = max(1, want // (W * H)); cw, ch = W, H
    paths, rays, dt = crop_run(cw, ch, spp)
    return dict(paths=paths, rays=rays, seconds=dt, cores=threads, sample=f"centred {cw}x{ch} crop of the {W}x{H} film, {spp} spp, all bounces")


def run_reference(args, rank, world, emit):
    if rank != 0:
        return
    pkg = load_pkg()
    defaults = pkg.config_defaults(CONFIG)
    W, H = defaults["width"], defaults["height"]
    steps, warm = args.steps, args.warmup
    # each step = a bounded sample; the whole run should end within a few minutes
    per_step = max(2.0, min(20.0, 150.0 / max(steps + warm, 1)))
    out = None
    try:
        runs = []
        for i in range(warm + steps):
            r = cpu_reference_run(defaults, args.level, per_step, args.cpu_threads)
            if r is None:
                break
            if i >= warm:
                runs.append(r)
        if runs:
            rays = sum(r["rays"] for r in runs); paths = sum(r["paths"] for r in runs); secs = sum(r["seconds"] for r in runs)
            value = rays / secs / 1e6
            out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
                   "ms_per_step": secs / len(runs) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                   "data": "synthetic", "config": {"workload": defaults["name"], "width": W, "height": H, "max_depth": defaults["max_depth"],
                                                   "triangles": 1310720, "step": runs[-1]["sample"]},
                   "spp_per_s": paths / secs / (W * H), "Mpaths_per_s": paths / secs / 1e6,
                   "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": runs[-1]["cores"], "kind": "reference", "sample": runs[-1]["sample"]},
                   "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    except Exception as e:  # the oracle always exists in a built tree; say why if not
        out = {"impl": "reference", "unavailable": f"{type(e).__name__}: {e}"}
    if out is None:
        out = {"impl": "reference", "unavailable": "oracle/_ref/libagpt_ref.so not built (needs /root/reference at build time)"}
    emit(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--spp-per-step", type=int, default=16)
    ap.add_argument("--level", type=int, default=0, help="icosphere subdivision override (tests); 0 = the configuration's own")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    # stdout carries exactly one JSON line: park fd 1 on stderr until then (NCCL / libraries may print)
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world, emit)
        return

    import torch
    import torch.distributed as dist
    import numpy as np

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference for the CPU reference)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    pkg = load_pkg()
    defaults = pkg.config_defaults(CONFIG)
    W = args.width or defaults["width"]; H = args.height or defaults["height"]
    depth, depth_arg = defaults["max_depth"], defaults["depth_arg"]
    spp_step = args.spp_per_step

    scene = pkg.HostScene(CONFIG, args.level)
    counts = scene.counts()
    ctx = pkg.Context(local_rank)
    stream = torch.cuda.Stream(device=local_rank)     # kernels, events and NCCL all on this one stream
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    scene.upload(ctx)
    ctx.set_film(W, H)
    accum = torch.zeros((H, W, 4), dtype=torch.float32, device=f"cuda:{local_rank}")   # the float4 accumulator NCCL reduces
    ctx.set_accum_dev(accum.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # samples are split by index across ranks: rank g renders s = g (mod world)
    def step(k, flags=0):
        pkg.multigpu.render_sharded(ctx, k * spp_step * world, spp_step * world, depth, depth_arg, rank, world, flags)

    for k in range(args.warmup):
        step(k)
    if world > 1:
        pkg.multigpu.allreduce_accumulator(accum)         # warm-up of the collective too (NCCL sets its channels up on first use)
    accum.zero_()
    ctx.reset_stats()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for k in range(args.steps):
        step(args.warmup + k)
    if world > 1:
        pkg.multigpu.allreduce_accumulator(accum)         # float4[W*H] accumulators -> final framebuffer (NVLink)
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    st = ctx.stats()
    t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    tot = torch.tensor([float(st.rays), float(st.paths), float(st.kernel_launches), float(st.rays_reference_equivalent),
                        float(st.rays_closest), float(st.rays_shadow), float(st.rays_mis), float(st.rays_mis_culled), float(st.rays_tail_culled)],
                       dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_max = float(t.item())
    rays, paths, launches, rays_ref_eq, r_closest, r_shadow, r_mis, r_mis_culled, r_tail_culled = [float(v) for v in tot.tolist()]
    value = rays / ms_max / 1e3

    # ---- end-to-end through the reference-facing host API (host buffers in and out) -------
    tracer = pkg.HostTracer(depth, local_rank)
    host_acc, _film_owner = pkg.pinned_film(W, H)          # page-locked host film, as the host mirror's Accumulator allocates it
    tracer.render(scene, W, H, host_acc, 0, spp_step, depth_arg)      # uploads the scene, sizes the wavefront state, warms up
    e2e_steps = max(1, min(args.steps, 3))
    tctx_stats0 = None
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        tracer.render(scene, W, H, host_acc, rank + (1 + k) * spp_step * world, spp_step, depth_arg)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = (rays / args.steps) * e2e_steps / float(e2e_t.item()) / 1e6    # same rays per step as the timed region
    tracer.close()

    # ---- roofline of the dominant kernel + counters (rank 0, extra untimed steps) --------
    roofline = None
    breakdown = None
    if rank == 0:
        ctx.reset_stats()
        step(args.warmup + args.steps, pkg.FLAG_TIMING)
        tm = ctx.stats()
        ctx.reset_stats()
        step(args.warmup + args.steps, pkg.FLAG_COUNTERS)
        cn = ctx.stats()
        peak, peak_src = measured_peak_gbs()
        bytes_closest = cn.algorithmic_bytes(0)
        achieved = bytes_closest / (tm.ms_trace_closest * 1e-3) / 1e9
        rays_closest = cn.rays_closest + cn.rays_mis
        roofline = {"bound": "hbm", "kernel": "k_trace_closest", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": None, "peak_source": peak_src, "launches_per_step": tm.launches_closest,
                    "avg_launch_ms": tm.ms_trace_closest / max(tm.launches_closest, 1),
                    "algorithmic_bytes_per_launch": bytes_closest / max(cn.launches_closest, 1),
                    "bytes_per_ray": bytes_closest / max(rays_closest, 1),
                    "interior_visits_per_ray": cn.node_visits[0] / max(rays_closest, 1),
                    "tri_tests_per_ray": cn.tri_tests[0] / max(rays_closest, 1),
                    "analytic_tests_per_ray": cn.analytic_tests[0] / max(rays_closest, 1)}
        breakdown = {"ms_step": tm.ms_render, "ms_trace_closest": tm.ms_trace_closest, "ms_trace_any": tm.ms_trace_any, "ms_shade": tm.ms_shade,
                     "ms_other": tm.ms_other, "waves": tm.waves, "rays_per_path": cn.rays / max(cn.paths, 1)}
        try:
            with open(os.path.join(ROOT, "profiles", "traffic_r1.json")) as f:
                per_ray = json.load(f).get("k_trace_closest_dram_bytes_per_ray")
                # ncu --set full capture (profiles/), scaled to this run's average launch
                roofline["traffic"] = per_ray * rays_closest / max(cn.launches_closest, 1) if per_ray else None
        except Exception:
            pass

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = cpu_reference_run(defaults, args.level, args.cpu_seconds, args.cpu_threads)
            if r:
                cpu = {"value": r["rays"] / r["seconds"] / 1e6, "unit": "Mrays/s", "cores": r["cores"], "kind": "reference", "sample": r["sample"],
                       "Mpaths_per_s": r["paths"] / r["seconds"] / 1e6, "seconds": r["seconds"]}
        except Exception as e:
            cpu = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic",
               "config": {"workload": defaults["name"], "width": W, "height": H, "max_depth": depth, "spp_per_step_per_gpu": spp_step,
                          "triangles": counts["tris"], "bvh_nodes": counts["nodes"], "primitives": counts["prims"], "lights": counts["lights"],
                          "scene_bytes": counts["bytes"], "parallelism": f"sample-index split x{world}, one NCCL all-reduce of the float4 accumulator",
                          "l2": "inputs larger than L2 (scene + wavefront state >> 126 MB); no flush needed"},
               "spp_per_s": paths / (W * H) / (ms_max * 1e-3), "Mpaths_per_s": paths / ms_max / 1e3,
               # value counts rays actually traced through the scene.  The reference also traces rays whose
               # outcome cannot matter (MIS rays that miss their light's sphere, the discarded ray at MaxDepth);
               # the B200 path proves them useless and skips them, so spp/s is the like-for-like speed and
               # ref_equivalent_Mrays_per_s is the throughput in the reference's own ray accounting.
               "ref_equivalent_Mrays_per_s": rays_ref_eq / ms_max / 1e3,
               "rays": {"closest_path": r_closest, "shadow": r_shadow, "mis_traced": r_mis, "mis_culled_exactly": r_mis_culled,
                        "tail_culled_exactly": r_tail_culled},
               "clocks": clocks,
               "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": W * H * 16 + 76, "d2h_bytes_per_step": W * H * 16,
                       "api": "CudaPathTracer::Render over host Accumulator buffers", "steps": e2e_steps},
               "gpu_launches": int(launches),
               "roofline": roofline, "breakdown": breakdown, "cpu_baseline": cpu}
        emit(out)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
