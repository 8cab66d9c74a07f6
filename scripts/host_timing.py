import importlib, time, sys, os
sys.path.insert(0, '/root/repo')
gp = importlib.import_module("go-pbrt_b200"); P = gp.pbrt
import torch
dev = P.Device(0)
scene, integ = gp.scenes.config2()
g = P.GpuScene(dev, scene)
film_dev = torch.zeros(1920*1080*4, dtype=torch.float64, device="cuda")
for i in range(3): P.Render(g, integ, 1, mode=1, device_film=film_dev.data_ptr())
os.environ["GOPBRT_HOST_TIMING"] = "1"
for i in range(2):
    t0 = time.time(); st = P.Render(g, integ, 1, mode=1, device_film=film_dev.data_ptr()); t1 = time.time()
    print("device-film call wall %.1f ms, library device %.1f ms" % ((t1 - t0) * 1e3, st["ms_total"]), flush=True)
