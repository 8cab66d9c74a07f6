"""Multi-GPU behind the C ABI: the film reduce (gopbrt_comm_*, GOPBRT_FLAG_REDUCE_FILM) and the one-process /
N-device form of pbrt.Render (gopbrt_multi_*).  The one-GPU cases run on every box (a world-1 communicator still goes through
libnccl); the N-device cases need >= 2 GPUs in the process and are skipped otherwise (scripts/multi_check.py runs them under
`gpurun --gpus N`)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gpu_count():
    cudart = C.CDLL("libcudart.so.12")
    n = C.c_int(0)
    cudart.cudaGetDeviceCount(C.byref(n))
    return n.value


def test_world1_communicator_reduce_is_identity(gp):
    # a communicator of one rank: ncclReduce in place leaves the film as rendered; the NCCL binding itself is exercised
    P, abi = gp.pbrt, gp.abi
    d = P.Device(0)
    d.comm_init(P.Device.comm_unique_id(), 0, 1)
    scene, integ = gp.scenes.config1(W=96, H=54)
    g = P.GpuScene(d, scene)
    P.Render(g, integ, 1, mode=abi.MODE_FAST)
    plain = integ.GetCamera().GetFilm().pixels.copy()
    st = P.Render(g, integ, 1, mode=abi.MODE_FAST, flags=abi.FLAG_REDUCE_FILM)
    assert np.array_equal(integ.GetCamera().GetFilm().pixels, plain)
    assert st["ms_reduce"] > 0
    with pytest.raises(RuntimeError, match="rank/world"):  # render options must agree with the communicator
        P.Render(g, integ, 1, mode=abi.MODE_FAST, flags=abi.FLAG_REDUCE_FILM, rank=0, world=2)
    g.close()
    d.close()


def test_reduce_flag_without_communicator_fails(gp, dev):
    P, abi = gp.pbrt, gp.abi
    scene, integ = gp.scenes.config1(W=32, H=18)
    g = P.GpuScene(dev, scene)
    with pytest.raises(RuntimeError, match="communicator"):
        P.Render(g, integ, 1, flags=abi.FLAG_REDUCE_FILM)
    g.close()


def test_multi_one_device_equals_single_context(gp, dev):
    P, abi = gp.pbrt, gp.abi
    scene, integ = gp.scenes.mixed_test_scene(60, seed=3), gp.scenes.test_integrator(80, 48, spp=(3, 3), maxDepth=5)
    g = P.GpuScene(dev, scene)
    st = P.Render(g, integ, 1, mode=abi.MODE_FAST)
    single = integ.GetCamera().GetFilm().pixels.copy()
    g.close()
    m = P.MultiDevice(1)
    mg = P.MultiGpuScene(m, scene)
    mst = P.RenderMulti(mg, integ, 1, mode=abi.MODE_FAST)
    assert np.array_equal(integ.GetCamera().GetFilm().pixels, single)
    for k in ("camera_rays", "closest_rays", "shadow_rays"):
        assert mst[k] == st[k]
    mg.close()
    m.close()


@pytest.mark.parametrize("mode", ["fast", "strict"])
def test_multi_n_devices_sum_to_the_single_gpu_film(gp, dev, mode):
    n = min(_gpu_count(), 8)
    if n < 2:
        pytest.skip("needs >= 2 GPUs in one process")
    P, abi = gp.pbrt, gp.abi
    md = abi.MODE_FAST if mode == "fast" else abi.MODE_STRICT
    scene, integ = gp.scenes.config2(W=160, H=90)
    g = P.GpuScene(dev, scene)
    st = P.Render(g, integ, 1, mode=md)
    single = integ.GetCamera().GetFilm().pixels.copy()
    g.close()
    m = P.MultiDevice(n)
    mg = P.MultiGpuScene(m, scene)
    for _ in range(2):
        mst = P.RenderMulti(mg, integ, 1, mode=md)
        film = integ.GetCamera().GetFilm().pixels
        assert np.array_equal(film[..., 3], single[..., 3])            # weights: integers, exact in any order
        assert np.allclose(film, single, rtol=1e-12, atol=1e-300)      # radiance: the same addends in another order
        for k in ("camera_rays", "closest_rays", "shadow_rays"):
            assert mst[k] == st[k], k
        assert mst["ms_reduce"] > 0
    mg.close()
    m.close()


def test_plain_c_caller_runs_the_whole_boundary(gp, tmp_path):
    # tests/cpp/abi_smoke.c: scene_create -> trace_closest / trace_any -> render -> cancel semantics -> gopbrt_multi_* from C99
    import subprocess
    from test_abi import _build_abi_smoke
    gp.abi.load()
    out = subprocess.run([_build_abi_smoke(tmp_path)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "abi_smoke: ok" in out.stdout, out.stdout[-3000:]
