"""dump GPU-vs-oracle closest-hit mismatches of the mixed test scene to gpurun_out/mismatch.npz (debug aid)"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
gp = importlib.import_module("go-pbrt_b200")
from oracle_lib import OracleScene
dev = gp.pbrt.Device(0)
scene = gp.scenes.mixed_test_scene(300)
g = gp.pbrt.GpuScene(dev, scene)
s = OracleScene(scene, 1)
rng = np.random.default_rng(3)
n = 200000
o = rng.uniform(-15, 15, size=(n, 3)); o[:, 1] = rng.uniform(0, 12, size=n)
d = rng.normal(size=(n, 3))
tm = np.where(rng.uniform(size=n) < 0.3, rng.uniform(1, 30, size=n), np.inf)
G = g.Intersect(o, d, tm); O = s.intersect(o, d, tm)
bad = np.nonzero((G[0] != O[0]) | (G[1] != O[1]))[0]
print("mismatches", len(bad))
for i in bad[:20]:
    print(i, "gpu", G[0][i], repr(G[1][i]), "oracle", O[0][i], repr(O[1][i]), "tmax", tm[i])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez(os.path.join(ROOT, "gpurun_out", "mismatch.npz"), o=o[bad], d=d[bad], tm=tm[bad], gprim=G[0][bad], gt=G[1][bad], oprim=O[0][bad], ot=O[1][bad])
