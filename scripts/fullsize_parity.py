"""BASELINE-size parity: the full 1920x1080 films of configs 1 and 2 (STRICT, tileSize 1) and config 2 in FAST mode (one
lane per pixel), GPU vs the oracle on all host cores, every pixel compared bit for bit.  Writes gpurun_out/fullsize_parity.json."""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
gp = importlib.import_module("go-pbrt_b200"); P = gp.pbrt
from oracle_lib import OracleScene
dev = P.Device(0)
out = {}
threads = os.cpu_count() or 8
for name, cfg, kw in (("config1 strict", "config1", dict()), ("config2 strict", "config2", dict()), ("config2 fast", "config2", dict(mode=gp.abi.MODE_FAST, groups=1))):
    scene, integ = getattr(gp.scenes, cfg)()
    g = P.GpuScene(dev, scene)
    st = P.Render(g, integ, 1, **kw)
    film = integ.GetCamera().GetFilm().pixels.copy()
    g.close()
    o = OracleScene(scene, 1)
    t0 = time.time()
    ofilm, ost = o.render(integ, 1, mode=kw.get("mode", 0), threads=threads)
    osec = time.time() - t0
    o.close()
    bad = np.any(film != ofilm, axis=2)
    rel = np.abs(film[..., :3] - ofilm[..., :3]).sum() / max(np.abs(ofilm[..., :3]).sum(), 1e-300)
    out[name] = dict(pixels=int(bad.size), pixels_differing=int(bad.sum()), weights_equal=bool(np.array_equal(film[..., 3], ofilm[..., 3])),
                     rays_gpu=[st["camera_rays"], st["closest_rays"], st["shadow_rays"]], rays_oracle=[ost["camera_rays"], ost["closest_rays"], ost["shadow_rays"]],
                     relative_l1_difference=float(rel), gpu_ms=st["ms_total"], oracle_seconds=osec, oracle_threads=threads,
                     differing_pixels_xy=[[int(x), int(y)] for y, x in np.argwhere(bad)[:16]])
    print(name, json.dumps(out[name]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fullsize_parity.json"), "w"), indent=1)
