"""The batched Aggregate.Intersect / IntersectP drop-ins against the independent plain-Python restatement: tests/golden/hits_golden.json
(tests/golden/make_hits_golden.py) holds, for camera rays and second-generation rays on four scenes, the primitive, tHit, hit point
and geometric normal, and the any-hit bit of shadow segments.  The oracle must answer identically — bit for bit — with the
reference-faithful BVH (accel 0), its own tree (1) and brute force (2).  (The CUDA path's gopbrt_trace_closest / _any are compared
with the oracle on far more rays in tests/test_gpu_parity.py and tests/test_gpu_configs.py.)  Nothing here reads /root/reference."""
import importlib.util
import json
import os

import numpy as np
import pytest

from oracle_lib import OracleScene

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_hits_golden", os.path.join(HERE, "golden", "make_hits_golden.py"))
X = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(X)

with open(os.path.join(HERE, "golden", "hits_golden.json")) as _f:
    RAW = json.load(_f)["cases"]
FH = float.fromhex


def _rays(rs):
    return (np.array([[FH(v) for v in r["o"]] for r in rs]), np.array([[FH(v) for v in r["d"]] for r in rs]),
            np.array([np.inf if r["tmax"] == "inf" else FH(r["tmax"]) for r in rs]))


def test_golden_file_covers_what_it_claims():
    assert sorted(RAW) == sorted(X.CASES)
    for name, c in RAW.items():
        hits = [h for h in c["hits"] if h]
        assert len(c["rays"]) >= 250 and len(hits) > 100 and len({h["prim"] for h in hits}) >= 5 and len(c["shadow_rays"]) > 100
    assert sum(RAW["partial"]["occluded"]) > 10 and sum(RAW["spheres"]["occluded"]) > 10


@pytest.mark.parametrize("name", sorted(X.CASES))
def test_generator_is_deterministic_and_matches_the_committed_file(gp, name):
    assert X.rays_and_answers(gp, name) == RAW[name]


@pytest.mark.parametrize("accel", [0, 1, 2])
@pytest.mark.parametrize("name", sorted(X.CASES))
def test_oracle_answers_the_independent_hit_records(gp, name, accel):
    c = RAW[name]
    scene, _ = X.scene_and_integrator(gp, name)
    o = OracleScene(scene, accel)
    prim, t, p, n = o.intersect(*_rays(c["rays"]))
    occluded = o.intersect_p(*_rays(c["shadow_rays"]))
    o.close()
    want = np.array([-1 if h is None else h["prim"] for h in c["hits"]])
    hit = want >= 0
    assert np.array_equal(prim, want)
    assert np.array_equal(t[hit], np.array([FH(h["t"]) for h in c["hits"] if h]))
    assert np.array_equal(p[hit], np.array([[FH(v) for v in h["p"]] for h in c["hits"] if h]))
    assert np.array_equal(n[hit], np.array([[FH(v) for v in h["n"]] for h in c["hits"] if h]))
    assert np.array_equal(occluded, np.array(c["occluded"]))
