"""per-iteration queue sizes and stage times of one frame (debug aid): GOPBRT_ITER_LOG"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["GOPBRT_ITER_LOG"] = os.path.join(ROOT, "gpurun_out", os.environ.get("ITER_LOG_NAME", "iter_log.csv"))
gp = importlib.import_module("go-pbrt_b200")
cfg = sys.argv[1] if len(sys.argv) > 1 else "config2"
scene, integ = getattr(gp.scenes, cfg)()
dev = gp.pbrt.Device(0); g = gp.pbrt.GpuScene(dev, scene)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
buf = None
import torch
film = torch.zeros(1920 * 1080 * 4, dtype=torch.float64, device="cuda")
for i in range(2):
    st = gp.pbrt.Render(g, integ, 1, flags=gp.abi.FLAG_TIME_KERNELS, device_film=film.data_ptr(), mode=int(os.environ.get("AB_MODE", "0")))
print({k: st[k] for k in ("ms_total", "ms_raygen", "ms_extend", "ms_shade", "ms_shadow", "iterations", "closest_rays", "shadow_rays")})
