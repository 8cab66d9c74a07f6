"""config 4 (10 M-triangle heightfield, 1080p) under the profiler / the iteration log.
  cfg4_profile.py frame [fast|strict]   one warm-up-free frame (for ncu: -k regex:k_trace -c 6 = extend+shadow of iterations 0..2)
  cfg4_profile.py iters [fast|strict]   per-iteration queue sizes + stage times -> gpurun_out/iter_cfg4_<mode>.csv, counters -> json"""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
what = sys.argv[1] if len(sys.argv) > 1 else "frame"
mode_s = sys.argv[2] if len(sys.argv) > 2 else "fast"
cfg = sys.argv[3] if len(sys.argv) > 3 else "config4"
OUT = os.path.join(ROOT, "gpurun_out"); os.makedirs(OUT, exist_ok=True)
if what == "frame":
    os.environ["GOPBRT_NO_GRAPH"] = "1"
gp = importlib.import_module("go-pbrt_b200"); P = gp.pbrt
mode = gp.abi.MODE_FAST if mode_s == "fast" else gp.abi.MODE_STRICT
scene, integ = getattr(gp.scenes, cfg)()
dev = P.Device(0); g = P.GpuScene(dev, scene)
import torch
film = torch.zeros(1920 * 1080 * 4, dtype=torch.float64, device="cuda")
if what == "frame":
    st = P.Render(g, integ, 1, mode=mode, device_film=film.data_ptr())
    print(json.dumps({k: st[k] for k in ("ms_total", "iterations", "lanes", "closest_rays", "shadow_rays")}))
else:
    P.Render(g, integ, 1, mode=mode, device_film=film.data_ptr())
    c = P.Render(g, integ, 1, mode=mode, flags=gp.abi.FLAG_COUNT_TRAVERSAL, device_film=film.data_ptr())
    os.environ["GOPBRT_ITER_LOG"] = os.path.join(OUT, f"iter_{cfg}_{mode_s}.csv")
    st = P.Render(g, integ, 1, mode=mode, flags=gp.abi.FLAG_TIME_KERNELS, device_film=film.data_ptr())
    del os.environ["GOPBRT_ITER_LOG"]
    t = P.Render(g, integ, 1, mode=mode, flags=gp.abi.FLAG_TIME_KERNELS, device_film=film.data_ptr())
    p = P.Render(g, integ, 1, mode=mode, device_film=film.data_ptr())
    out = dict(config=cfg, mode=mode_s, counters=c, timed=t, plain_ms=p["ms_total"])
    json.dump(out, open(os.path.join(OUT, f"counters_{cfg}_{mode_s}.json"), "w"), indent=1)
    print(json.dumps({k: t[k] for k in ("ms_total", "ms_raygen", "ms_extend", "ms_shade", "ms_shadow", "iterations", "lanes")}), p["ms_total"])
