"""One rank's share of the headline frame on one GPU (what a rank of an N-GPU run executes, without the reduce): frame time and stage times.
Usage: python scripts/rank_share.py <world> [rank]"""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gp = importlib.import_module("go-pbrt_b200")
P = gp.pbrt
world = int(sys.argv[1]); rank = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = P.Device(0)
scene, integ = gp.scenes.config2()
g = P.GpuScene(dev, scene)
kw = dict(mode=gp.abi.MODE_FAST, rank=rank, world=world)
for _ in range(3):
    p = P.Render(g, integ, 1, **kw)
t = P.Render(g, integ, 1, flags=gp.abi.FLAG_TIME_KERNELS, **kw)
print(json.dumps(dict(world=world, rank=rank, ms_plain=round(p["ms_total"], 3), lanes=p["lanes"], iters=p["iterations"],
                      stage={k[3:]: round(t[k], 3) for k in ("ms_raygen", "ms_extend", "ms_shade", "ms_shadow", "ms_film")})))
g.close()
