"""Host phase timing of gopbrt_scene_create on the 10 M-triangle mesh (GOPBRT_HOST_TIMING; the `timing` variant build adds the
builder's own phases).  Usage: GOPBRT_LIB=.../variants/lib_timing.so python scripts/scene_build_timing.py [config4|config3]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["GOPBRT_HOST_TIMING"] = "1"
gp = importlib.import_module("go-pbrt_b200")
P = gp.pbrt
cfg = sys.argv[1] if len(sys.argv) > 1 else "config4"
t0 = time.time(); scene, integ = getattr(gp.scenes, cfg)(); t1 = time.time()
d = scene.desc(); t2 = time.time()
dev = P.Device(0)
for k in range(2):
    t3 = time.time(); g = P.GpuScene(dev, scene); t4 = time.time()
    print(f"[python] scene objects {t1 - t0:.2f} s, desc {t2 - t1:.2f} s, gopbrt_scene_create {t4 - t3:.3f} s", flush=True)
    g.close()
