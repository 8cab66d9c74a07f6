"""N > 1 path on CPU: two gloo ranks render their share of the frame (tiles t % world == rank in STRICT mode, samples
s % world == rank in FAST mode — the partition gopbrt_render_options describes), sum the films with ONE reduce, and the
result must equal the single-rank film up to summation order.  The renderer here is the oracle; the collective plumbing
(torch.distributed reduce on a W*H*4 float64 film) is the one bench.py uses with NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, mode, direct, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib
    gp = importlib.import_module("go-pbrt_b200")
    from oracle_lib import OracleScene
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene, integ = gp.scenes.config1(W=64, H=36)
    if direct:  # integrator.DirectLighting, every light sampled (UniformSampleAll)
        integ = gp.pbrt.NewDirectLighting(gp.pbrt.UniformSampleAll, 5, integ.GetCamera(), integ.GetSampler(), None)
    o = OracleScene(scene, 1)
    film, st = o.render(integ, 1, mode=mode, rank=rank, world=world, threads=2)
    t = torch.from_numpy(film.copy())
    paths = torch.tensor([st["camera_rays"]], dtype=torch.int64)
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    dist.all_reduce(paths, op=dist.ReduceOp.SUM)
    if rank == 0:
        single, st1 = o.render(integ, 1, mode=mode, threads=2)
        q.put((t.numpy(), single, int(paths[0]), st1["camera_rays"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,direct", [(0, False), (1, False), (1, True)])
def test_two_rank_gloo_partition_and_reduce(mode, direct):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, direct, q)) for r in range(2)]
    for p in procs:
        p.start()
    summed, single, paths, paths1 = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert paths == paths1 == 64 * 36 * 15
    assert np.array_equal(summed[..., 3], single[..., 3])      # filterWeightSum is integer-valued: exact
    assert np.allclose(summed, single, rtol=1e-12, atol=0)      # radiance: equal up to summation order
