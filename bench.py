#!/usr/bin/env python
"""bench.py — Mrays/s and 1080p frame time of the go-pbrt hot path on N B200s, beside the CPU reference.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one frame: one pass of pbrt.Render's hot path (raygen+sampler -> extend -> shade -> shadow -> film) over the
workload BASELINE.json's metric is quoted on (configs[1]): the procedural Cornell-box-style room of triangles, Path
integrator, Stratified 8x8 ("64 spp" = 63 effective samples), 1920x1080.  Sampler mode: FAST by default at
every N (counter-based streams per (pixel, sample): the north star's split by sample index across GPUs, lane groups inside
one GPU); --mode strict = one reference RNG stream per pixel (identical per-pixel sample sequences to
pbrt.Render(..., tileSize=1)), also reported beside the headline at N = 1 ("strict_mode").

value  = Mrays/s (closest-hit + any-hit queries the reference semantics require; the always-discarded MIS ray of
         EstimateDirect is neither traced nor counted), scene resident in HBM, film left on the device; device time
         from CUDA events on the library's stream, max over ranks.
e2e    = the same metric through the public C-ABI call with a HOST film buffer (gopbrt_render): per-step descriptors
         host->device, W*H*4 float64 film device->host.  N > 1: every rank makes the same call with
         GOPBRT_FLAG_REDUCE_FILM — the library sums the ranks' films with ONE ncclReduce on its own stream (the
         communicator is the library's: gopbrt_comm_init_rank) and rank 0 reads the host film.  torch.distributed is
         plumbing only (rendezvous, the 128-byte communicator id, the max-over-ranks of the timings); it moves no film.
config5 = BASELINE configs[4] (4K, 1023 spp, 10 M triangles, samples split over the N GPUs + the NCCL film reduce):
         1 warm-up + 2 timed frames through the same e2e call, reported as a sub-record at every N.
The reference arm (--impl reference) times the oracle — the C++ restatement of the Go renderer; Go itself cannot be
built or run in this image — on the box's host cores over a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "Mrays/s (closest-hit + any-hit) at 1080p, Path integrator, fixed spp"
WORKLOADS = {
    "config2": dict(name="config2: procedural Cornell-box-style room of triangles + matte/glass spheres, disk area light, "
                         "Path maxDepth 10, Stratified 8x8 (63 effective spp), 1920x1080", W=1920, H=1080, spp=(8, 8)),
    "config3": dict(name="config3: random 100k-sphere field, mixed materials, 8 sphere area lights + distant light, Path maxDepth 10, "
                         "Stratified 16x16 (255 effective spp), 1920x1080", W=1920, H=1080, spp=(16, 16)),
    "config4": dict(name="config4: 10 M-triangle heightfield (2237x2237 grid), Path maxDepth 10, Stratified 4x4 (15 effective spp), "
                         "1920x1080", W=1920, H=1080, spp=(4, 4)),
    "config5": dict(name="config5: 10 M-triangle heightfield, Path maxDepth 10, Stratified 32x32 (1023 effective spp), 3840x2160, "
                         "samples split across the GPUs + one NCCL film reduce", W=3840, H=2160, spp=(32, 32)),
    "config1": dict(name="config1: README sphere scene (internal/render/server.go), Path maxDepth 10, Stratified 4x4 "
                         "(15 effective spp), 1920x1080", W=1920, H=1080, spp=(4, 4)),
}
# algorithmic bytes per unit: 32 B per BVH node visited; per shape test 72 B (triangle: 9 f64), 32 B
# (translation-only sphere: centre + radius) or 224 B (general sphere/disk: two 3x4 f64 matrices + 4 scalars);
# per closest-hit ray 72 B of ray I/O (read o,d,tMax = 56 B, write t + primitive = 16 B), per any-hit ray 60 B
B_NODE, B_TRI, B_SPH, B_GEN, B_IO_CLOSEST, B_IO_ANY = 32, 72, 32, 224, 72, 60


def build_scene(gp, args):
    wl = WORKLOADS[args.config]
    W, H = args.width or wl["W"], args.height or wl["H"]
    scene, integ = getattr(gp.scenes, args.config)(W=W, H=H, spp=wl["spp"])
    return wl, W, H, scene, integ


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 8:
                self.rows.append(f)

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = sorted(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if r[4 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def bench_config(wl, W, H, tile, mode_name):
    """`config` of the JSON line: names the workload, identical in both arms (ours and --impl reference)"""
    return {"workload": wl["name"], "resolution": [W, H], "tile_size": tile, "sampler_mode": mode_name,
            "spp_effective": wl["spp"][0] * wl["spp"][1] - 1}


def cpu_reference(gp, args, steps, warmup, threads=None):
    """the oracle (C++ restatement of the Go renderer) on the host cores: same scene, same spp, same sampler mode, 1/16 of the
    pixels"""
    from oracle_lib import OracleScene
    wl = WORKLOADS[args.config]
    W, H = (args.width or wl["W"]) // 4, (args.height or wl["H"]) // 4
    scene, integ = getattr(gp.scenes, args.config)(W=W, H=H, spp=wl["spp"])
    threads = threads or os.cpu_count() or 1
    mode = gp.abi.MODE_STRICT if args.mode == "strict" else gp.abi.MODE_FAST
    o = OracleScene(scene, 1)
    secs, rays = [], 0
    for i in range(warmup + steps):
        film, st = o.render(integ, args.tile, mode=mode, threads=threads, deterministic=False)
        if i >= warmup:
            secs.append(st["seconds"])
            rays = st["closest_rays"] + st["shadow_rays"]
    o.close()
    t = sum(secs) / max(1, len(secs))
    return {"value": rays / t / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
            "sample": f"same scene, spp and sampler mode ({args.mode}) at {W}x{H} (1/16 of the frame's pixels), tileSize {args.tile}, {rays} rays per step, "
                      f"{t:.2f} s per step; C++ restatement of the Go renderer (no Go toolchain in this image)",
            "ms_per_step": t * 1e3, "rays_per_step": rays}


def algorithmic_bytes(c):
    """bytes the extend kernel's work is made of, from an instrumented frame's counters"""
    ext_rays = c["closest_rays"] - c["root_culled_rays"]
    by = (B_NODE * (c["nodes_visited"] - c["root_culled_rays"]) + B_TRI * c["tests_triangle"] + B_SPH * c["tests_sphere_fast"] +
          B_GEN * c["tests_general"] + B_IO_CLOSEST * ext_rays)
    return by, ext_rays


def deep_bvh_roofline(gp, g, integ, peak, mode, mode_name, build_s):
    """extend-kernel roofline on BASELINE configs[3] (10 M-triangle heightfield, 1080p, 15 spp): the scene whose node and
    triangle records (0.28 + 0.8 GB) do not fit L2, i.e. where the node-fetch roofline is an HBM roofline.
    Same sampler mode as the headline."""
    P, abi = gp.pbrt, gp.abi
    P.Render(g, integ, 1, mode=mode)
    c = P.Render(g, integ, 1, mode=mode, flags=abi.FLAG_COUNT_TRAVERSAL)
    ts = [P.Render(g, integ, 1, mode=mode, flags=abi.FLAG_TIME_KERNELS) for _ in range(3)]
    plain = [P.Render(g, integ, 1, mode=mode) for _ in range(3)]
    by, ext_rays = algorithmic_bytes(c)
    ms_ext = sum(t["ms_extend"] for t in ts) / len(ts)
    ms_frame = sum(t["ms_total"] for t in plain) / len(plain)
    rays = plain[0]["closest_rays"] + plain[0]["shadow_rays"]
    achieved = by / (ms_ext / 1e3) / 1e9
    return {"workload": "config4: 10 M-triangle heightfield (2237x2237 grid), Path, Stratified 4x4 (15 effective spp), 1920x1080",
            "sampler_mode": mode_name, "bound": "hbm", "kernel": "k_extend", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "bytes_per_frame": by, "ms_extend_per_frame": ms_ext, "extend_launches": ts[0]["extend_launches"],
            "per_ray": {"node_records_tested": (c["nodes_visited"] - c["root_culled_rays"]) / max(1, ext_rays),
                        "shape_tests": c["prim_tests"] / max(1, ext_rays)},
            "stage_ms_per_frame": {k: sum(t[k] for t in ts) / len(ts) for k in ("ms_raygen", "ms_extend", "ms_shade", "ms_shadow", "ms_film")},
            "frame_time_ms": ms_frame, "mrays_per_s": rays / ms_frame / 1e3, "bvh_nodes": ts[0]["bvh_nodes"], "bvh_depth": ts[0]["bvh_depth"],
            "scene_create_s": build_s}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="config2", choices=list(WORKLOADS))
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--tile", type=int, default=1)
    ap.add_argument("--mode", default="fast", choices=["strict", "fast"],
                    help="fast (default): counter-based sampler, a pixel's samples are independent -> lane groups on one GPU, "
                         "split by sample index across GPUs (north star); strict: the unmodified reference's per-pixel RNG "
                         "streams (tileSize 1), tiles split across GPUs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--min-warmup", type=int, default=3, help="lower bound on warm-up frames (supplementary long-frame workloads only)")
    ap.add_argument("--no-deep-bvh", action="store_true", help="skip the config-4 (10 M triangles) extend-kernel roofline")
    ap.add_argument("--no-config5", action="store_true", help="skip the config-5 (4K, 1023 spp, 10 M triangles) sub-record")
    ap.add_argument("--no-extra", action="store_true", help="skip the config-1 / config-3 lines under `extra`")
    ap.add_argument("--profile", action="store_true",
                    help="profiling run for ncu: exactly --warmup warm-up frames (0 allowed), --steps timed frames, "
                         "no instrumented pass, no e2e pass, no CPU baseline; the JSON line is NOT a bench value")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    gp = importlib.import_module("go-pbrt_b200")
    wl = WORKLOADS[args.config]
    W, H = args.width or wl["W"], args.height or wl["H"]
    config = bench_config(wl, W, H, args.tile, args.mode)

    if args.impl == "reference":
        if rank != 0:
            return 0
        nw = max(0, min(args.warmup, 1))
        cb = cpu_reference(gp, args, args.steps, nw)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "Mrays/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": nw, "warmup_requested": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    P, abi = gp.pbrt, gp.abi
    dev = P.Device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        # the library's own communicator for the film reduce: rank 0 makes the id, torch.distributed only carries its 128 bytes
        idt = torch.zeros(abi.COMM_ID_BYTES, dtype=torch.uint8, device=f"cuda:{local_rank}")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(P.Device.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, src=0)
        dev.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
    mode = abi.MODE_STRICT if args.mode == "strict" else abi.MODE_FAST
    RED = abi.FLAG_REDUCE_FILM if world > 1 else 0
    scene, integ = getattr(gp.scenes, args.config)(W=W, H=H, spp=wl["spp"])
    t0 = time.time()
    g = P.GpuScene(dev, scene)
    scene_create_s = time.time() - t0
    film_dev = torch.zeros(H * W * 4, dtype=torch.float64, device=f"cuda:{local_rank}")
    film_host = torch.empty(H * W * 4, dtype=torch.float64).pin_memory()
    film_host_np = film_host.numpy()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=f"cuda:{local_rank}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def sum_over_ranks(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=f"cuda:{local_rank}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.tolist()

    def step_device(flags, scene_h=None, integrator=None, film_ptr=None):
        """one frame, film stays on the device (N > 1: summed onto rank 0's by the library's ncclReduce); returns (library stats,
        device ms = wavefront + film merge + reduce, CUDA events on the library stream)"""
        st = P.Render(scene_h or g, integrator or integ, args.tile, mode=mode, rank=rank, world=world, flags=flags | RED,
                      device_film=film_ptr or film_dev.data_ptr())
        return st, st["ms_total"] + st["ms_reduce"]

    for _ in range(args.warmup if args.profile else max(args.min_warmup, args.warmup)):
        step_device(0)
    if args.profile:
        for _ in range(args.steps):
            st, ms = step_device(abi.FLAG_TIME_KERNELS)
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": ms, "stage_ms": {k: st[k] for k in ("ms_raygen", "ms_extend", "ms_shade", "ms_shadow", "ms_film")}}))
        g.close()
        dev.close()
        return 0
    # instrumented (untimed) passes of the same frame: (1) V = BVH node records tested, T = shape tests by kind;
    # (2) per-stage CUDA-event times (events between the stage launches keep the wavefront off its CUDA-graph path, so
    # these frames are a few percent slower than the timed ones: they give the extend kernel's own duration, not `value`)
    cst, _ = step_device(abi.FLAG_COUNT_TRAVERSAL)
    stage_stats = [step_device(abi.FLAG_TIME_KERNELS)[0] for _ in range(min(2, args.steps))]

    barrier()
    launches0 = dev.launches()
    per_step_ms, stats = [], []
    # clocks / throttle reasons are sampled over the device-timed loop (CUDA events: host-side stalls do not enter `value`).  The
    # sampler is an nvidia-smi process; NVML queries hold a driver lock for milliseconds at a time (tens on a multi-GPU box) and
    # CUDA API calls queue behind it, so it is started before and fully stopped after that loop and never runs during the
    # wall-clock e2e loops below.
    clocks = ClockSampler(local_rank)
    clocks.__enter__()
    time.sleep(0.3)
    barrier()
    wall0 = time.time()
    for _ in range(args.steps):
        st, ms = step_device(0)
        per_step_ms.append(ms)
        stats.append(st)
    barrier()
    wall = time.time() - wall0
    clocks.__exit__(None, None, None)
    time.sleep(1.0)
    launches = dev.launches() - launches0
    rays_rank = sum(s["closest_rays"] + s["shadow_rays"] for s in stats)
    ms_rank = sum(per_step_ms)

    # ---- e2e: the public C-ABI call with a HOST film buffer, every rank the same call (N > 1: + GOPBRT_FLAG_REDUCE_FILM)
    desc_bytes = sum(C.sizeof(t) for t in (abi.Camera, abi.Sampler, abi.Integrator, abi.Film, abi.RenderOptions))

    def step_e2e(out, scene_h=None, integrator=None):
        return P.Render(scene_h or g, integrator or integ, args.tile, mode=mode, rank=rank, world=world, flags=RED, out=out if rank == 0 else None)

    step_e2e(film_host_np)  # one untimed call through the host-film entry point: it allocates that path's device film on first use
    barrier()
    e0 = time.time()
    e2e_parts = {"ms_device": 0.0, "ms_reduce": 0.0, "ms_download": 0.0}
    e2e_walls = []
    for _ in range(args.steps):
        w0 = time.time()
        est = step_e2e(film_host_np)  # gopbrt_render: film D2H (into pinned host memory) inside the call
        e2e_walls.append((time.time() - w0) * 1e3)
        e2e_parts["ms_device"] += est["ms_total"] / args.steps
        e2e_parts["ms_reduce"] += est["ms_reduce"] / args.steps
        e2e_parts["ms_download"] += est.get("ms_download", 0.0) / args.steps
    barrier()
    e2e_s = time.time() - e0
    if os.environ.get("GOPBRT_BENCH_DEBUG"):
        print("[bench] e2e per-call wall ms:", [round(x, 1) for x in e2e_walls], "loop total", round(e2e_s * 1e3, 1), file=sys.stderr)
    # the same call into a PAGEABLE host buffer (a Go []float64 is pageable): one warm call, then a few timed ones
    film_pageable = np.empty(H * W * 4, dtype=np.float64) if rank == 0 else None
    step_e2e(film_pageable)
    barrier()
    p0 = time.time()
    n_pageable = min(3, args.steps)
    for _ in range(n_pageable):
        step_e2e(film_pageable)
    barrier()
    e2e_pageable_s = (time.time() - p0) / n_pageable

    ms_total, e2e_total, wall_max, e2e_pageable = max_over_ranks([ms_rank, e2e_s, wall, e2e_pageable_s])
    rays_total, launches_total, paths_total = sum_over_ranks([float(rays_rank), float(launches), float(stats[0]["camera_rays"])])
    iters_all = max_over_ranks([float(stats[0]["iterations"]), float(stats[0]["lanes"])])

    K = args.steps
    line = None
    if rank == 0:
        value = rays_total / (ms_total / 1e3) / 1e6
        e2e_value = rays_total / e2e_total / 1e6
        # ---- roofline of the dominant kernel (k_extend: closest-hit traversal), rank 0, per launch
        KS = len(stage_stats)
        s_ext = sum(s["ms_extend"] for s in stage_stats)
        n_ext = sum(s["extend_launches"] for s in stage_stats)
        # rays the extend kernel itself processed: camera rays that miss the BVH root are answered inside raygen
        bytes_frame, ext_rays = algorithmic_bytes(cst)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = bytes_frame * KS / (s_ext / 1e3) / 1e9 if s_ext > 0 else 0.0
        small = stats[0]["bvh_nodes"] <= 200  # the flat aggregate (<= 64 primitives): its table lives in shared memory
        # DRAM bytes per launch of the extend kernel.  NOT measured in this run (a timed run carries no profiler): one
        # `ncu --set full` capture of an extend launch of this workload (profiles/r02_traffic.json: dram__bytes_read.sum +
        # dram__bytes_write.sum per ray of that launch), scaled to this run's average rays per launch; null for other workloads.
        traffic, traffic_source = None, "no ncu capture of this workload under profiles/"
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            if tj.get("workload") == args.config:
                traffic = tj["dram_bytes_per_ray"] * ext_rays / max(1, n_ext / KS)
                traffic_source = ("profiles/r02_traffic.json: %.1f DRAM bytes per ray in one ncu --set full capture of an extend launch "
                                  "(%d rays), scaled to this run's rays per launch - not measured in this run" % (tj["dram_bytes_per_ray"], tj["rays_in_launch"]))
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": "k_extend (closest-hit traversal, conservative f32 node boxes + fp64 own-bound / EFloat / watertight shape tests)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6.65 TB/s (of fallback)",
                "traffic": traffic,
                "traffic_source": traffic_source,
                "limiter": ("issue / SIMT divergence: the scene's %d-entry bound table is shared-memory resident and every ray tests every entry "
                            "(V = table size), so the algorithmic node bytes never reach DRAM; see roofline_deep_bvh for the HBM-resident tree" % cst["bvh_nodes"]
                            if small else "node-fetch latency (L2/HBM) and SIMT divergence"),
                "bytes_per_launch": bytes_frame * KS / max(1, n_ext), "ms_per_launch": s_ext / max(1, n_ext), "launches": n_ext,
                "measured_over": "%d instrumented frame(s) of the same workload (CUDA events around every stage launch)" % KS,
                "extend_rays_per_step": ext_rays,
                "per_ray": {"nodes_visited": (cst["nodes_visited"] - cst["root_culled_rays"]) / max(1, ext_rays),
                            "shape_tests": cst["prim_tests"] / max(1, ext_rays)},
                "stage_ms_per_step": {k: sum(s[k] for s in stage_stats) / KS for k in ("ms_raygen", "ms_extend", "ms_shade", "ms_shadow", "ms_film")}}
        # ---- the kernel that dominates THIS frame.  On the headline workload (36 primitives: the aggregate is a shared-memory table)
        # that is the shade stage, whose bytes all come from and go to HBM: per shaded lane it reads the lane's PathRec (64 B:
        # throughput, sampler state), the refraction scale in its RadRec sector (32 B), the RayRec (64 B) and the queue entry
        # (4 B) and writes the sampler-state sector back (32 B); a lane that continues also writes its throughput sector, a new
        # RayRec and its queue entry (32 + 64 + 4 B), a lane with a light sample a ShadowRec and its queue entry (96 + 4 B) —
        # DESIGN.md §4.
        # (ms_shade = k_split_hits + the shade-class kernels of every iteration.)  The extend kernel stays on the line as
        # roofline_extend: its SURVEY §8d "algorithmic bytes" are table bytes that never leave the SM here, so their rate is not
        # a fraction of an HBM roof — for the HBM-resident tree see roofline_deep_bvh.
        s_shade = sum(s["ms_shade"] for s in stage_stats)
        st0 = stage_stats[0]
        if s_shade > s_ext and st0.get("shaded_lanes", 0) > 0:
            roof_ext = dict(roof)
            if small:
                roof_ext["bound"] = "issue / SIMT divergence (on-chip table)"
                roof_ext["algorithmic_GBps"] = roof_ext.pop("achieved")
                roof_ext.pop("frac", None)
            cont_lanes = st0["closest_rays"] - st0["camera_rays"]  # every closest-hit query that is not a camera ray was spawned by a shade lane
            n_iter = st0["iterations"]
            n_cls = max(1, round((st0["launches"] - 3) / max(1, n_iter)) - 5)  # shade-class launches per iteration (launches per iteration = 5 + classes)
            by_shade = st0["shaded_lanes"] * (96 + 64 + 4 + 32) + cont_lanes * (32 + 64 + 4) + st0["shadow_rays"] * (96 + 4)
            n_sh = n_iter * n_cls * KS
            ach = by_shade * KS / (s_shade / 1e3) / 1e9
            traffic_sh, traffic_sh_src = None, "no ncu capture of this workload under profiles/"
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic_shade.json")))
                if tj.get("workload") == args.config:
                    traffic_sh = tj["dram_bytes_per_shaded_lane"] * st0["shaded_lanes"] / max(1, n_iter * n_cls)
                    traffic_sh_src = ("profiles/r02_traffic_shade.json: %.1f DRAM bytes per shaded lane over the shade-class launches of one "
                                      "wavefront iteration in one ncu --set full capture, scaled to this run's lanes per launch - not measured in this run"
                                      % tj["dram_bytes_per_shaded_lane"])
            except Exception:
                pass
            roof = {"bound": "hbm", "kernel": "k_shade (+ k_split_hits): one Path.Li loop body per lane on its HBM-resident lane records",
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "peak_source": roof_ext["peak_source"], "traffic": traffic_sh, "traffic_source": traffic_sh_src,
                    "limiter": "dependent float64 chains at 16 resident warps per SM (128-168 registers): issue slots 33 % busy, long-scoreboard 5 of 12 stall cycles per issue (profiles/r02b_ncu_stage_kernels.md)",
                    "bytes_per_launch": by_shade * KS / max(1, n_sh), "ms_per_launch": s_shade / max(1, n_sh), "launches": n_sh,
                    "bytes_per_unit": "196 B per shaded lane + 100 B per continuing lane + 100 B per shadow ray",
                    "units_per_step": {"shaded_lanes": st0["shaded_lanes"], "continuing_lanes": cont_lanes, "shadow_rays": st0["shadow_rays"]},
                    "measured_over": roof_ext["measured_over"],
                    "stage_ms_per_step": roof_ext["stage_ms_per_step"]}
        else:
            roof_ext = None
        lane_bytes = 64 + 96 + (64 + 32 + 32) + 36 + 1 + (0 if mode == abi.MODE_FAST else 4 * wl["spp"][0] * wl["spp"][1] * 8)
        line = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": max(args.min_warmup, args.warmup),
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config,
                "run": {"paths_per_step": int(paths_total), "rays_per_step": rays_total / K, "lanes_per_gpu": int(iters_all[1]),
                        "wavefront_iterations": int(iters_all[0]),
                        "partition": ("tiles t %% %d == rank" % world) if mode == abi.MODE_STRICT else "samples s %% %d == rank" % world,
                        "film_reduce": "ncclReduce(float64, sum) inside gopbrt_render on the library stream" if world > 1 else None,
                        "l2_policy": "per-lane path/ray/shadow state of %d lanes is %.1f GB >> 126 MB L2 (inputs larger than L2)" % (
                            stats[0]["lanes"], stats[0]["lanes"] * lane_bytes / 1e9),
                        "scene_create_s": scene_create_s, "bvh_nodes": stats[0]["bvh_nodes"], "bvh_depth": stats[0]["bvh_depth"],
                        "frame_time_ms": ms_total / K, "frame_time_e2e_ms": e2e_total / K * 1e3,
                        "wall_ms_per_step": wall_max / K * 1e3},
                "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": desc_bytes, "d2h_bytes_per_step": H * W * 4 * 8,
                        "ms_per_step": e2e_total / K * 1e3, "host_buffer": "pinned", **e2e_parts,
                        "pageable_host_buffer": {"value": rays_total / K / e2e_pageable / 1e6, "ms_per_step": e2e_pageable * 1e3, "steps": n_pageable}},
                "gpu_launches": int(launches_total), "clocks": clocks.summary(), "roofline": roof,
                **({"roofline_extend": roof_ext} if roof_ext else {}),
                "reference_panics": {"radiance_gt10": stats[0]["radiance_gt10"], "efloat_panics": stats[0]["efloat_panics"],
                                     "nan_samples": stats[0]["nan_samples"]}}
    if world == 1 and mode == abi.MODE_FAST:
        # the same frame in STRICT mode (the unmodified reference's per-pixel sample sequences, pbrt.Render(..., tileSize=1))
        sm = [P.Render(g, integ, args.tile, mode=abi.MODE_STRICT, device_film=film_dev.data_ptr()) for _ in range(K + 1)][1:]
        line["strict_mode"] = {"value": sum(x["closest_rays"] + x["shadow_rays"] for x in sm) / sum(x["ms_total"] for x in sm) / 1e3,
                               "unit": "Mrays/s", "ms_per_step": sum(x["ms_total"] for x in sm) / K,
                               "wavefront_iterations": sm[0]["iterations"], "lanes": sm[0]["lanes"],
                               "note": "bit-exact against the oracle's STRICT mode (tests/test_gpu_parity.py); FAST is bit-exact against the "
                                       "oracle's FAST mode and statistically equivalent to STRICT (SURVEY §8d tier 3, same test file); both "
                                       "modes reproduce, bit for bit, the films of an independent plain-Python restatement of the Go source "
                                       "on configs 1 and 2 (tests/test_config{1,2}_golden.py, test_fast_golden.py, test_baseline_spp_golden.py)"}
    g.close()
    del film_dev, film_host, film_host_np

    # ---- the deep-BVH workloads: configs[3] (extend-kernel roofline where the tree lives in HBM, N = 1) and configs[4]
    # (4K, 1023 spp, samples split over the GPUs + the NCCL film reduce, every N)
    want_deep = world == 1 and not args.no_deep_bvh
    want_c5 = not args.no_config5 and args.config == "config2"
    if want_deep or want_c5:
        mesh_scene, integ4 = gp.scenes.config4()
        mesh_scene.desc()  # the host's flat arrays (the Python stand-in for the Go exporter): not part of gopbrt_scene_create
        t0 = time.time()
        g4 = P.GpuScene(dev, mesh_scene)  # gopbrt_scene_create: raw upload + on-device BVH build + records
        build_s = time.time() - t0
        if want_deep:
            line["roofline_deep_bvh"] = deep_bvh_roofline(gp, g4, integ4, peak, mode, args.mode, build_s)
        if want_c5:
            w5 = WORKLOADS["config5"]
            integ5 = gp.scenes.config4_integrator(w5["W"], w5["H"], w5["spp"])
            film5 = np.empty(w5["W"] * w5["H"] * 4, dtype=np.float64) if rank == 0 else None
            step_e2e(film5, g4, integ5)  # warm-up frame (allocates the lane state and the 4K device film)
            barrier()
            c0 = time.time()
            c5 = [step_e2e(film5, g4, integ5) for _ in range(2)]
            barrier()
            c5_s = (time.time() - c0) / 2
            (c5_wall, c5_dev, c5_red) = max_over_ranks([c5_s, sum(x["ms_total"] for x in c5) / 2, sum(x["ms_reduce"] for x in c5) / 2])
            (c5_rays,) = sum_over_ranks([float(sum(x["closest_rays"] + x["shadow_rays"] for x in c5) / 2)])
            if rank == 0:
                line["config5"] = {"workload": w5["name"], "n_gpus": world, "warmup": 1, "steps": 2, "value": c5_rays / c5_wall / 1e6, "unit": "Mrays/s",
                                   "frame_time_ms": c5_wall * 1e3, "ms_device_max": c5_dev, "ms_reduce_max": c5_red,
                                   "rays_per_step": c5_rays, "lanes_per_gpu": c5[0]["lanes"], "wavefront_iterations": c5[0]["iterations"],
                                   "timed": "wall clock between barriers around gopbrt_render with a (pageable) HOST film on rank 0: descriptors H2D, "
                                            "wavefront, film merge, ncclReduce, 265 MB film D2H",
                                   "scene_create_s": build_s}
        g4.close()
    if rank == 0 and world == 1 and not args.no_extra and args.config == "config2":
        # BASELINE configs[0] and configs[2] through the same device-film call (1 warm-up + 2 timed frames each)
        line["extra"] = {}
        for name in ("config1", "config3"):
            w = WORKLOADS[name]
            sc_x, integ_x = getattr(gp.scenes, name)()
            gx = P.GpuScene(dev, sc_x)
            fx = torch.zeros(w["H"] * w["W"] * 4, dtype=torch.float64, device=f"cuda:{local_rank}")
            rec = {}
            for mname, m in (("fast", abi.MODE_FAST), ("strict", abi.MODE_STRICT)):
                rs = [P.Render(gx, integ_x, 1, mode=m, device_film=fx.data_ptr()) for _ in range(3)][1:]
                rec[mname] = {"value": sum(x["closest_rays"] + x["shadow_rays"] for x in rs) / sum(x["ms_total"] for x in rs) / 1e3, "unit": "Mrays/s",
                              "frame_time_ms": sum(x["ms_total"] for x in rs) / 2, "wavefront_iterations": rs[0]["iterations"]}
            ts = P.Render(gx, integ_x, 1, mode=mode, flags=abi.FLAG_TIME_KERNELS, device_film=fx.data_ptr())
            cc = P.Render(gx, integ_x, 1, mode=mode, flags=abi.FLAG_COUNT_TRAVERSAL, device_film=fx.data_ptr())
            by, er = algorithmic_bytes(cc)
            rec["extend"] = {"ms_per_frame": ts["ms_extend"], "algorithmic_GBps": by / (ts["ms_extend"] / 1e3) / 1e9 if ts["ms_extend"] > 0 else None,
                             "node_records_per_ray": (cc["nodes_visited"] - cc["root_culled_rays"]) / max(1, er), "shape_tests_per_ray": cc["prim_tests"] / max(1, er)}
            rec["workload"] = w["name"]
            rec["bvh_nodes"] = ts["bvh_nodes"]
            line["extra"][name] = rec
            gx.close()
            del fx
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference(gp, args, 1, 0)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    dev.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
