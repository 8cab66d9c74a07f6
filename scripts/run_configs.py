"""BASELINE configs 1/3/4 on the GPU: scene build + upload time, a 1080p frame, and oracle parity on sampled primary
rays and on a small film (tileSize 1).  Writes gpurun_out/configs.json."""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
gp = importlib.import_module("go-pbrt_b200")
from oracle_lib import OracleScene, camera_rays
P = gp.pbrt
dev = P.Device(0)
out = {}
which = sys.argv[1:] or ["config1", "config3", "config4"]
small = {"config1": dict(W=160, H=90), "config3": dict(W=96, H=54, spp=(3, 3)), "config4": dict(W=96, H=54)}
for name in which:
    t0 = time.time(); scene, integ = getattr(gp.scenes, name)(); t_host = time.time() - t0
    t0 = time.time(); scene.desc(); t_desc = time.time() - t0
    t0 = time.time(); g = P.GpuScene(dev, scene); t_up = time.time() - t0
    st = P.Render(g, integ, 1, flags=gp.abi.FLAG_COUNT_TRAVERSAL)
    st = P.Render(g, integ, 1, flags=gp.abi.FLAG_TIME_KERNELS)
    film = integ.GetCamera().GetFilm().pixels
    rays = st["closest_rays"] + st["shadow_rays"]
    r = dict(host_build_s=t_host, desc_s=t_desc, scene_create_s=t_up, ms_frame=st["ms_total"], mrays_s=rays / st["ms_total"] / 1e3,
             rays=rays, paths=st["camera_rays"], iterations=st["iterations"], bvh_nodes=st["bvh_nodes"], bvh_depth=st["bvh_depth"],
             gt10=st["radiance_gt10"], efloat_panics=st["efloat_panics"], nan=st["nan_samples"], stack_overflows=st["stack_overflows"],
             stage_ms={k: st[k] for k in ("ms_raygen", "ms_extend", "ms_shade", "ms_shadow")}, finite=bool(np.isfinite(film).all()),
             mean_xyz=[float(x) for x in film[..., :3].mean(axis=(0, 1))])
    # parity: sampled primary rays against the oracle (its own tree + the parity-spec own-bound test)
    t0 = time.time(); o = OracleScene(scene, 1); r["oracle_build_s"] = time.time() - t0
    rng = np.random.default_rng(1)
    xs, ys = rng.integers(0, 1920, 20000), rng.integers(0, 1080, 20000)
    ro, rd = camera_rays(integ, xs, ys)
    G = g.Intersect(ro, rd); O = o.intersect(ro, rd, threads=16)
    r["primary_hitmiss_equal"] = bool(np.array_equal(G[0] >= 0, O[0] >= 0))
    r["primary_prim_mismatch"] = int(np.count_nonzero(G[0] != O[0]))
    r["primary_t_mismatch"] = int(np.count_nonzero(G[1] != O[1]))
    r["primary_hits"] = int(np.count_nonzero(O[0] >= 0))
    g.close()
    # parity: small film
    scene2, integ2 = getattr(gp.scenes, name)(**small[name]) if name != "config4" else (scene, gp.scenes.config4.__wrapped__ if False else None)
    if name == "config4":
        # reuse the 10M-triangle scene with a small film
        cam = gp.scenes._camera((0.0, 15.0, 80.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 50.0, 96, 54)
        integ2 = P.NewPath(10, cam, P.NewStratified(4, 4, False, 4), None, 1, P.Uniform); scene2 = scene
        o2 = o
    else:
        o.close(); o2 = OracleScene(scene2, 1)
    g2 = P.GpuScene(dev, scene2)
    st2 = P.Render(g2, integ2, 1); f2 = integ2.GetCamera().GetFilm().pixels.copy()
    of, ost = o2.render(integ2, 1, threads=16)
    bad = int(np.count_nonzero(np.any(f2 != of, axis=2)))
    r["small_film_pixels_differing"] = bad; r["small_film_pixels"] = int(f2.shape[0] * f2.shape[1])
    r["small_film_rays_gpu_oracle"] = [st2["closest_rays"] + st2["shadow_rays"], ost["closest_rays"] + ost["shadow_rays"]]
    g2.close(); o2.close()
    out[name] = r
    print(name, json.dumps(r), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
