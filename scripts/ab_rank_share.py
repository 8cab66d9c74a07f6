"""One rank's share of the config-2 frame on one GPU (rank 0 of `world`), for several lane-group counts: the per-GPU frame time
a world-N run would see, without N GPUs.  Usage: ab_rank_share.py world g1 g2 ..."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gp = importlib.import_module("go-pbrt_b200"); P = gp.pbrt
world = int(sys.argv[1])
dev = P.Device(0)
scene, integ = gp.scenes.config2()
g = P.GpuScene(dev, scene)
for grp in [int(x) for x in sys.argv[2:]]:
    kw = dict(mode=gp.abi.MODE_FAST, rank=0, world=world, groups=grp)
    for _ in range(2):
        P.Render(g, integ, 1, **kw)
    t = P.Render(g, integ, 1, flags=gp.abi.FLAG_TIME_KERNELS, **kw)
    p = [P.Render(g, integ, 1, **kw)["ms_total"] for _ in range(3)]
    print(json.dumps(dict(world=world, groups=grp, ms_plain=round(min(p), 2), iters=t["iterations"], lanes=t["lanes"],
                          stage={k[3:]: round(t[k], 2) for k in ("ms_raygen", "ms_extend", "ms_shade", "ms_shadow", "ms_film")})), flush=True)
