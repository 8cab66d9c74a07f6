"""Root-causes a GPU-vs-oracle ray-count mismatch of a STRICT (tileSize 1) frame: binary search over the tile partition
(tiles t % world == rank on both sides) down to the single tile whose counts differ, then the per-iteration log of that
one lane on the GPU (GOPBRT_ITER_LOG) beside the oracle's event trace (ORACLE_TRACE) of the same tile.
Usage: find_ray_mismatch.py [config2]  -> gpurun_out/ray_mismatch.json (+ .gpu.csv / .oracle.txt)"""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
gp = importlib.import_module("go-pbrt_b200"); P = gp.pbrt
from oracle_lib import OracleScene
OUT = os.path.join(ROOT, "gpurun_out"); os.makedirs(OUT, exist_ok=True)
cfg = sys.argv[1] if len(sys.argv) > 1 else "config2"
threads = os.cpu_count() or 8
scene, integ = getattr(gp.scenes, cfg)()
dev = P.Device(0); g = P.GpuScene(dev, scene); o = OracleScene(scene, 1)
W, H = 1920, 1080

def both(rank, world, thr=threads):
    st = P.Render(g, integ, 1, rank=rank, world=world)
    gf = integ.GetCamera().GetFilm().pixels.copy()
    of, ost = o.render(integ, 1, rank=rank, world=world, threads=thr)
    return (st["closest_rays"], st["shadow_rays"]), (ost["closest_rays"], ost["shadow_rays"]), int(np.count_nonzero(np.any(gf != of, axis=2)))

log = []
t0 = time.time()
gc, oc, bad = both(0, 1)
log.append(dict(rank=0, world=1, gpu=gc, oracle=oc, pixels_differing=bad))
print(log[-1], flush=True)
res = dict(config=cfg, levels=log)
if gc != oc:
    rank, world = 0, 1
    while world < W * H:
        world *= 2
        cand = [rank, rank + world // 2]
        found = None
        for r in cand:
            gc, oc, bad = both(r, world)
            log.append(dict(rank=r, world=world, gpu=gc, oracle=oc, pixels_differing=bad))
            print(log[-1], flush=True)
            if gc != oc:
                found = r
                break
        if found is None:
            res["error"] = "mismatch vanished when the partition was refined"
            break
        rank = found
    else:
        tile = rank
        res["tile"] = tile; res["pixel_xy"] = [tile % W, tile // W]
        # one lane alone: per-iteration GPU log and the oracle's event trace
        os.environ["GOPBRT_ITER_LOG"] = os.path.join(OUT, "ray_mismatch.gpu.csv")
        st = P.Render(g, integ, 1, rank=rank, world=world, flags=gp.abi.FLAG_TIME_KERNELS)
        del os.environ["GOPBRT_ITER_LOG"]
        os.environ["ORACLE_TRACE"] = os.path.join(OUT, "ray_mismatch.oracle.txt")
        of, ost = o.render(integ, 1, rank=rank, world=world, threads=1)
        del os.environ["ORACLE_TRACE"]
        res["single_tile"] = dict(gpu=[st["closest_rays"], st["shadow_rays"]], oracle=[ost["closest_rays"], ost["shadow_rays"]])
        # align: k-th oracle C event <-> k-th GPU iteration with an extend ray; compare "a shadow ray followed"
        import csv
        rows = [r for r in csv.DictReader(open(os.path.join(OUT, "ray_mismatch.gpu.csv"))) if int(r["extend_rays"]) > 0]
        ev = []
        for line in open(os.path.join(OUT, "ray_mismatch.oracle.txt")):
            if line.startswith("C "):
                ev.append(dict(line=line.strip(), shadow=0, E=None))
            elif line.startswith("  E") and ev:
                ev[-1]["E"] = line.strip(); ev[-1]["shadow"] = int(line.strip().rsplit("shadow=", 1)[1])
        res["events_gpu"], res["events_oracle"] = len(rows), len(ev)
        for k, (r, e) in enumerate(zip(rows, ev)):
            if int(r["shadow_rays"]) != e["shadow"]:
                res["first_divergence"] = dict(event=k, gpu_shadow=int(r["shadow_rays"]), oracle=e)
                break
res["seconds"] = time.time() - t0
json.dump(res, open(os.path.join(OUT, "ray_mismatch.json"), "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if k != "levels"}))
