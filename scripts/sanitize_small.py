"""Small renders touching every kernel variant, for `compute-sanitizer --tool memcheck|racecheck python scripts/sanitize_small.py`."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gp = importlib.import_module("go-pbrt_b200"); P = gp.pbrt
dev = P.Device(0)
scene, integ = gp.scenes.config2(W=48, H=27, spp=(3, 3))
g = P.GpuScene(dev, scene)
for kw in (dict(), dict(mode=gp.abi.MODE_FAST), dict(mode=gp.abi.MODE_FAST, groups=3, max_lanes=1000), dict(flags=gp.abi.FLAG_COUNT_TRAVERSAL | gp.abi.FLAG_TIME_KERNELS),
           dict(flags=gp.abi.FLAG_TAIL)):
    st = P.Render(g, integ, 1, **kw)
    print("path", kw, st["closest_rays"], st["shadow_rays"], flush=True)
for strat in (P.UniformSampleOne, P.UniformSampleAll):
    dl = P.NewDirectLighting(strat, 5, integ.GetCamera(), integ.GetSampler(), None)
    st = P.Render(g, dl, 1)
    print("direct", strat, st["closest_rays"], st["shadow_rays"], flush=True)
g.close()
scene = gp.scenes.mixed_test_scene(60, seed=3)
integ = gp.scenes.test_integrator(40, 30, spp=(2, 2), maxDepth=5)
g = P.GpuScene(dev, scene)
st = P.Render(g, integ, 8)
dl = P.NewDirectLighting(P.UniformSampleAll, 5, integ.GetCamera(), integ.GetSampler(), None)
st = P.Render(g, dl, 4, mode=gp.abi.MODE_FAST)
rng = np.random.default_rng(1)
o = rng.uniform(-15, 15, size=(5000, 3)); d = rng.normal(size=(5000, 3))
prim, t, p, n = g.Intersect(o, d)
occ = g.IntersectP(o, d)
print("mixed ok", int((prim >= 0).sum()), int(occ.sum()), flush=True)
g.close(); dev.close()
