"""Verification build of the interval-free sphere path (GOPBRT_LIB=.../variants/lib_chk.so, built with -DGP_CHECK_FAST_SPHERE): every
full-sphere test runs BOTH paths; a disagreement (hit / miss, root, tHit) or an interval wider than the bound the fast path
assumes raises the efloat_panics counter.  Renders configs 1-3 and the mixed scenes and prints the counters (all must be 0)."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gp = importlib.import_module("go-pbrt_b200"); P = gp.pbrt
dev = P.Device(0)
cases = [("config1", gp.scenes.config1()), ("config2", gp.scenes.config2()), ("config3 (64 spp)", gp.scenes.config3(spp=(8, 8))),
         ("mixed 200", (gp.scenes.mixed_test_scene(200, seed=7), gp.scenes.test_integrator(640, 360, spp=(4, 4), maxDepth=8))),
         ("mixed 3000", (gp.scenes.mixed_test_scene(3000, seed=9), gp.scenes.test_integrator(640, 360, spp=(4, 4), maxDepth=8)))]
for name, (scene, integ) in cases:
    g = P.GpuScene(dev, scene)
    for mode in (gp.abi.MODE_FAST, gp.abi.MODE_STRICT):
        st = P.Render(g, integ, 1, mode=mode, flags=gp.abi.FLAG_COUNT_TRAVERSAL)
        print(json.dumps(dict(case=name, mode=mode, efloat_panics=st["efloat_panics"], sphere_tests=st["tests_sphere_fast"] + st["tests_general"] + st["shadow_tests_sphere_fast"] + st["shadow_tests_general"],
                              rays=st["closest_rays"] + st["shadow_rays"])), flush=True)
    g.close()
