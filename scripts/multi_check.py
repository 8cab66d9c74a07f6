"""One process, N GPUs through gopbrt_multi_* (run under `gpurun --gpus N`): parity of the reduced film against one GPU, then
frame times of config 2 (1080p) through gopbrt_multi_render with a host film.  Usage: multi_check.py [N] [config]"""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
gp = importlib.import_module("go-pbrt_b200")
P, abi = gp.pbrt, gp.abi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
config = sys.argv[2] if len(sys.argv) > 2 else "config2"
scene, integ = getattr(gp.scenes, config)(W=320, H=180)
d = P.Device(0)
g = P.GpuScene(d, scene)
P.Render(g, integ, 1, mode=abi.MODE_FAST)
single = integ.GetCamera().GetFilm().pixels.copy()
g.close(); d.close()
m = P.MultiDevice(n)
mg = P.MultiGpuScene(m, scene)
P.RenderMulti(mg, integ, 1)
film = integ.GetCamera().GetFilm().pixels
print(json.dumps({"n": n, "weights_equal": bool(np.array_equal(film[..., 3], single[..., 3])),
                  "max_rel": float(np.max(np.abs(film - single) / np.maximum(np.abs(single), 1e-300)))}), flush=True)
mg.close()
scene, integ = getattr(gp.scenes, config)()
t0 = time.time()
mg = P.MultiGpuScene(m, scene)
t_scene = time.time() - t0
for i in range(5):
    t0 = time.time()
    st = P.RenderMulti(mg, integ, 1)
    wall = time.time() - t0
    print(json.dumps({"n": n, "config": config, "frame": i, "wall_ms": round(wall * 1e3, 2), "ms_device_max": round(st["ms_total"], 2), "ms_reduce": round(st["ms_reduce"], 3),
                      "ms_download": round(st["ms_download"], 2), "mrays_e2e": round((st["closest_rays"] + st["shadow_rays"]) / wall / 1e6, 1), "iters": st["iterations"],
                      "scene_s": round(t_scene, 2)}), flush=True)
mg.close(); m.close()
