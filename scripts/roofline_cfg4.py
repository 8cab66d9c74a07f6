"""config 4 (10M-triangle heightfield): extend-kernel algorithmic bytes vs time (the HBM-relevant roofline)"""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gp = importlib.import_module("go-pbrt_b200")
P = gp.pbrt
name = sys.argv[1] if len(sys.argv) > 1 else "config4"
scene, integ = getattr(gp.scenes, name)()
dev = P.Device(0); g = P.GpuScene(dev, scene)
P.Render(g, integ, 1)
c = P.Render(g, integ, 1, flags=gp.abi.FLAG_COUNT_TRAVERSAL)
t = P.Render(g, integ, 1, flags=gp.abi.FLAG_TIME_KERNELS)
t = P.Render(g, integ, 1, flags=gp.abi.FLAG_TIME_KERNELS)
by = 32 * c["nodes_visited"] + 72 * c["tests_triangle"] + 32 * c["tests_sphere_fast"] + 224 * c["tests_general"] + 72 * c["closest_rays"]
sby = 32 * c["shadow_nodes_visited"] + 72 * c["shadow_tests_triangle"] + 32 * c["shadow_tests_sphere_fast"] + 224 * c["shadow_tests_general"] + 60 * c["shadow_rays"]
out = dict(config=name, closest_rays=c["closest_rays"], shadow_rays=c["shadow_rays"], V_per_ray=c["nodes_visited"] / c["closest_rays"],
           T_per_ray=c["prim_tests"] / c["closest_rays"], sV_per_ray=c["shadow_nodes_visited"] / max(1, c["shadow_rays"]),
           extend_bytes=by, ms_extend=t["ms_extend"], extend_GBps=by / t["ms_extend"] / 1e6, frac_of_6447=by / t["ms_extend"] / 1e6 / 6447.2,
           shadow_bytes=sby, ms_shadow=t["ms_shadow"], shadow_GBps=sby / max(t["ms_shadow"], 1e-9) / 1e6,
           ms_total=t["ms_total"], mrays=(t["closest_rays"] + t["shadow_rays"]) / t["ms_total"] / 1e3, bvh_nodes=t["bvh_nodes"], bvh_depth=t["bvh_depth"],
           stage={k: t[k] for k in ("ms_raygen", "ms_extend", "ms_shade", "ms_shadow")})
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"roofline_{name}.json"), "w"), indent=1)
