"""A/B timing of the traversal kernels: renders the named configs and prints per-stage CUDA-event times.
GOPBRT_LIB=<variant .so> selects the build that runs; AB_MODE=1 = FAST mode, AB_GROUPS, AB_FLAGS, AB_W/AB_H as in the code.  Usage: ab_trace.py tag config2 [config4 ...]"""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gp = importlib.import_module("go-pbrt_b200")
P = gp.pbrt
tag = sys.argv[1]
dev = P.Device(0)
for name in sys.argv[2:]:
    dims = {k: int(os.environ[e]) for k, e in (('W', 'AB_W'), ('H', 'AB_H')) if e in os.environ}
    scene, integ = getattr(gp.scenes, name)(**dims)
    g = P.GpuScene(dev, scene)
    xf = int(os.environ.get("AB_FLAGS", "0"))
    kw = dict(mode=int(os.environ.get("AB_MODE", "0")), groups=int(os.environ.get("AB_GROUPS", "0")))
    P.Render(g, integ, 1, **kw)
    t = P.Render(g, integ, 1, flags=gp.abi.FLAG_TIME_KERNELS | xf, **kw)
    t = P.Render(g, integ, 1, flags=gp.abi.FLAG_TIME_KERNELS | xf, **kw)
    p0 = P.Render(g, integ, 1, flags=xf, **kw)
    p0 = P.Render(g, integ, 1, flags=xf, **kw)
    film = integ.GetCamera().GetFilm().pixels
    import hashlib
    out = dict(tag=tag, config=name, ms_plain=round(p0["ms_total"], 2), plain_iters=p0["iterations"], ms_total=round(t["ms_total"], 2),
               mrays=round((t["closest_rays"] + t["shadow_rays"]) / t["ms_total"] / 1e3, 1),
               stage={k[3:]: round(t[k], 2) for k in ("ms_raygen", "ms_extend", "ms_shade", "ms_shadow")},
               iters=t["iterations"], lanes=t["lanes"], ms_tail=round(t.get("ms_tail", 0), 2), paths=t["camera_rays"], rays=t["closest_rays"] + t["shadow_rays"], ovf=t["stack_overflows"], film_sha=hashlib.sha1(film.tobytes()).hexdigest()[:12])
    print(json.dumps(out), flush=True)
    g.close()
