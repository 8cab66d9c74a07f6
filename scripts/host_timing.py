"""Host-side phases of one gopbrt_render call (GOPBRT_HOST_TIMING): device-film and host-film entry points, config 2 at 1080p FAST."""
import importlib, time, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gp = importlib.import_module("go-pbrt_b200"); P = gp.pbrt
import torch
dev = P.Device(0)
scene, integ = gp.scenes.config2()
g = P.GpuScene(dev, scene)
film_dev = torch.zeros(1920 * 1080 * 4, dtype=torch.float64, device="cuda")
host = torch.empty(1920 * 1080 * 4, dtype=torch.float64).pin_memory().numpy()
for i in range(3):
    P.Render(g, integ, 1, mode=1, device_film=film_dev.data_ptr())
    P.Render(g, integ, 1, mode=1, out=host)
os.environ["GOPBRT_HOST_TIMING"] = "1"
for i in range(2):
    t0 = time.time(); st = P.Render(g, integ, 1, mode=1, device_film=film_dev.data_ptr()); t1 = time.time()
    print("device-film call wall %.1f ms, library device %.1f ms" % ((t1 - t0) * 1e3, st["ms_total"]), flush=True)
for i in range(3):
    t0 = time.time(); st = P.Render(g, integ, 1, mode=1, out=host); t1 = time.time()
    print("host-film call wall %.1f ms, library device %.1f ms, download %.2f ms" % ((t1 - t0) * 1e3, st["ms_total"], st["ms_download"]), flush=True)
