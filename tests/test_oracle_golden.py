"""Pins the oracle (oracle/*.h) and the host mirror against EVERY golden value the reference's own tests hold for the
hot path.  CPU only.  Each test cites the reference test it replays."""
import ctypes as C
import math

import numpy as np

from oracle_lib import OracleScene, vec


def test_efloat_add_golden(oracle):
    # pkg/efloat/efloat_test.go:9-13
    out = (C.c_double * 3)()
    oracle.oracle_kat_efloat_add(1.0, 0.0, 1.0, 0.0, out)
    assert list(out) == [2.0, 1.9999999999999998, 2.0000000000000004]


def test_offset_ray_origin_golden(oracle):
    # pkg/pbrt/ray_test.go:10-19 — pins the denormal MachineEpsilon and the x1024
    eps = oracle.oracle_kat_machine_epsilon()
    assert eps == 5e-324
    out = (C.c_double * 3)()
    oracle.oracle_kat_offset_ray_origin(vec(0, 0, 0), vec(eps, eps, eps), vec(1, 1, 1), vec(1, 1, 1), out)
    assert list(out) == [1.5183e-320] * 3
    g3 = oracle.oracle_kat_gamma(3.0)
    assert g3 == 1.5e-323 and 1 + 2 * g3 == 1.0


def test_matrix_inverse_golden(gp):
    # pkg/pbrt/transform_test.go:17-36
    P = gp.pbrt
    m = P.Matrix4x4([[1, 0, 0, 0], [0, 1, 1, 0], [0, 0, 1, 0], [0, 0, 2, 1]])
    assert m.Inverse().m == [[1, 0, 0, 0], [0, 1, -1, 0], [0, 0, 1, 0], [0, 0, -2, 1]]


def test_transform_point_golden(gp, oracle):
    # pkg/pbrt/transform_test.go:66-75
    P = gp.pbrt
    out = (C.c_double * 6)()
    t = P.NewTransform(P.Matrix4x4()).abi()
    oracle.oracle_kat_transform_point(C.byref(t), vec(0, 0, 0), vec(0, 0, 0), out)
    assert list(out) == [0.0] * 6
    t = P.Translate((5.0, 4.0, 3.0)).abi()
    oracle.oracle_kat_transform_point(C.byref(t), vec(0, 0, 0), vec(0, 0, 0), out)
    assert list(out)[:3] == [5.0, 4.0, 3.0]
    assert P.Translate((5.0, 4.0, 3.0)).TransformPoint((0.0, 0.0, 0.0)) == (5.0, 4.0, 3.0)


def test_transform_ray_golden(gp, oracle):
    # pkg/pbrt/transform_test.go:77-81 — RotateY(90)·Translate(5,4,3) on ray (0,0,0)->(1,0,0); needs Go's own cos
    P = gp.pbrt
    xf = P.RotateY(90).Mul(P.Translate((5.0, 4.0, 3.0)))
    out = (C.c_double * 6)()
    t = xf.abi()
    oracle.oracle_kat_transform_ray(C.byref(t), vec(0, 0, 0), vec(1, 0, 0), out)
    assert list(out) == [3.0000000000000004, 4.0, -5.0, 6.123233995736757e-17, 0.0, -1.0]
    assert math.cos(math.pi / 180 * 90) != 6.123233995736757e-17  # libm differs: why gomath exists
    assert gp.gomath.Cos(math.pi / 180 * 90) == 6.123233995736757e-17
    assert oracle.oracle_kat_trig(1, math.pi / 180 * 90, 0) == 6.123233995736757e-17


def test_visibility_tester_golden(oracle):
    # pkg/pbrt/light_test.go:10-44 — SpawnRayToInteraction (0,0,0)->(10,0,0): origin, dir (10,0,0), TMax 0.9999
    out = (C.c_double * 7)()
    oracle.oracle_kat_spawn_ray_to(vec(0, 0, 0), vec(10, 0, 0), out)
    assert list(out) == [0.0, 0.0, 0.0, 10.0, 0.0, 0.0, 0.9999]


def _bvh_scene(gp, max_prims):
    P = gp.pbrt
    return P.NewScene(P.NewBVH(gp.scenes.bvh_test_primitives(), max_prims, P.SplitSAH), [])


TEST_RAYS = [((0, 0, 0), (0, 0, 1.0)), ((0, 0, 0), (0, 0, -1.0)), ((0, 0, 500), (0, 0, -1.0)), ((10, 10, 500), (0, 0, -1.0))]


def test_bvh_intersect_golden(gp):
    # pkg/accelerator/bvh_test.go:43-98: NewBVH(primitives, 255, SplitSAH); expected primitive prim1, none, prim2, prim3
    for accel in (0, 1, 2):
        s = OracleScene(_bvh_scene(gp, 255), accel)
        prim, t, p, n = s.intersect([r[0] for r in TEST_RAYS], [r[1] for r in TEST_RAYS])
        assert list(prim) == [0, -1, 1, 2], accel
        assert list(t[[0, 2, 3]]) == [4.0, 489.0, 489.0]  # SURVEY App. F derived values
        s.close()


def test_bvh_intersectp_golden(gp):
    # pkg/accelerator/bvh_test.go:100-141 (maxPrims 2) and simple_test.go:59-108
    for accel in (0, 1, 2):
        s = OracleScene(_bvh_scene(gp, 2), accel)
        hit = s.intersect_p([r[0] for r in TEST_RAYS], [r[1] for r in TEST_RAYS])
        assert list(hit) == [True, False, True, True], accel
        s.close()


def test_simple_intersect_exact_point_golden(gp):
    # pkg/accelerator/simple_test.go:40-57 — assert.Equal on float64: hit point == (10,10,10) + normalize(1,1,1)
    s = OracleScene(_bvh_scene(gp, 255), 2)
    inv = 1.0 / math.sqrt(3.0)  # Normalize multiplies by 1/sqrt (xyz.go:587-595)
    d = (-1 * inv, -1 * inv, -1 * inv)
    rec = s.hit_record((15, 15, 15), d)
    assert rec is not None and int(rec[21]) == 2
    exp = 10 + 1 * inv
    assert rec[0:3] == [exp, exp, exp]
    assert rec[0] == 10.577350269189626 and rec[20] == 7.6602540378443855
    # miss (simple_test.go:51-57)
    assert s.hit_record((0, 0, 0), (0, 0, -1.0)) is None
    s.close()


def test_rng_variant_vectors(oracle, gp):
    # no reference golden exists for the RNG; these are the derived vectors of SURVEY App. F (Q29) and
    # cross-check the oracle against the independent host-side Python port in scenes.RNG
    exp = {0: [0xf0eefa37, 0xa7fcff1c, 0x58ff7ac6, 0xa7adaeec, 0x43800362],
           1: [0x00000003, 0x50c007f7, 0x80000006, 0x7200008c, 0x65fbf4ff],
           8159: [0xc5c2051d, 0xe0000008, 0x4b00fa58, 0x1a8001e9, 0x18a7cc01]}
    out = (C.c_uint32 * 5)()
    for seed, v in exp.items():
        oracle.oracle_kat_rng(seed, 1, 5, out)
        assert list(out) == v
        r = gp.scenes.RNG(seed)
        assert [r.UniformUInt32() for _ in range(5)] == v
    oracle.oracle_kat_rng(0, 0, 3, out)
    assert list(out)[:3] == [0x2a58a78d, 0x00000003, 0xdd0001f3]


def test_config1_camera_matrices(gp):
    # SURVEY App. F (derived; Q7/Q7b): the README camera is an off-axis pinhole with w == 1
    scene, integ = gp.scenes.config1()
    cam = integ.GetCamera()
    r2c = cam.RasterToCamera.Matrix.m
    assert abs(r2c[0][0] - 6.207049961428177e-4) < 1e-18 and abs(r2c[1][1] + 1.1034755486983426e-3) < 1e-18
    assert r2c[3] == [0.0, 0.0, 0.0, 1.0] and r2c[2][3] == 1.0
    c2w = cam.cameraToWorld.startTransform.Matrix.m
    assert np.allclose(c2w[0], [-0.9010475702906073, -0.2803300858899106, -0.3309506292762536, 150], rtol=0, atol=1e-15)
    assert cam.shutterClose == cam.shutterOpen == 0.0


def test_config1_reference_bvh_shape(gp):
    # faithful RecursiveBuild(SplitSAH, maxPrims=2) over the 23 README primitives -> 45 nodes
    scene, _ = gp.scenes.config1()
    s = OracleScene(scene, 0)
    assert s.lib.oracle_scene_bvh_nodes(s.h) == 45
    s.close()


def test_sampler_stratified_tables(oracle, gp):
    # Stratified(4,4,false,4): 2-D tables are all (0,0); 1-D tables are a permutation of (i+.5)/16; dims past
    # nSampledDimensions fall through to the RNG
    cfg = gp.abi.Sampler(0, 4, 4, 0, 4, 0)
    vals = set()
    for idx in range(1, 16):
        o1 = (C.c_double * 6)()
        o2 = (C.c_double * 12)()
        oracle.oracle_kat_sampler(C.byref(cfg), 5, 0, idx, 6, o1, o2)
        vals.add(o1[0])
        assert all(v == 0.0 for v in list(o2)[:8])
        assert all(0.0 <= v < 1.0 for v in o1)
        assert any(v != 0.0 for v in list(o2)[8:])
    assert vals <= {(i + 0.5) / 16 for i in range(16)} and len(vals) == 15


# ---- the reference's remaining unit tests for helpers the path is made of, replayed against the oracle's own functions
def _kat(oracle, fn, args, n_out=16):
    oracle.oracle_kat_eval.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_int]
    oracle.oracle_kat_eval.restype = C.c_int
    vin = (C.c_double * len(args))(*args)
    out = (C.c_double * n_out)()
    n = oracle.oracle_kat_eval(None, fn.encode(), vin, len(args), out, n_out)
    assert n >= 0, f"{fn}: rejected"
    return list(out)[:n]


def _xyz(oracle, op, a, b=(0, 0, 0), s=0.0):
    return _kat(oracle, "xyz_op", [op, *a, *b, s])


def test_xyz_golden(oracle):
    # pkg/geometry/xyz_test.go:9-162, every arithmetic case (String/Set/SetIndex are Go plumbing with no counterpart)
    assert _xyz(oracle, 0, (-1, -2, -3)) == [1, 2, 3]                       # Abs            :9-13
    assert _xyz(oracle, 1, (-1, -2, -3), (1, 2, 3)) == [14.0]               # AbsDot         :15-18
    assert _xyz(oracle, 2, (1, 2, 3), (1, 2, 3)) == [2, 4, 6]               # Add, AddAssign :20-29
    assert _xyz(oracle, 3, (1, 2, 3), s=1.0) == [2, 3, 4]                   # AddConst       :31-34
    assert _xyz(oracle, 4, (1, 2, 3), (1, 2, 4)) == [2, -1, 0]              # Cross          :36-39
    assert _xyz(oracle, 5, (1, 2, 3), (1, 2, 4)) == [1.0]                   # Distance       :41-44
    assert _xyz(oracle, 6, (1, 2, 3), (1, 2, 5)) == [4.0]                   # DistanceSquared:46-49
    assert _xyz(oracle, 7, (3, 9, 27), (3, 3, 3)) == [1, 3, 9]              # Div, DivAssign :51-60
    assert _xyz(oracle, 8, (3, 9, 27), s=3.0) == [1, 3, 9]                  # DivScalar      :62-65
    assert _xyz(oracle, 9, (3, 9, 27), (2, 4, 6)) == [204.0]                # Dot            :67-71
    assert _xyz(oracle, 16, (3, 9, 27), (3, 9, 27)) == [1] and _xyz(oracle, 16, (3, 9, 27), (0, 0, 0)) == [0]  # Equals / NotEquals :73-77,134-138
    assert [_xyz(oracle, 17, (2, 3, 4), s=k)[0] for k in range(3)] == [2, 3, 4]  # Index     :79-84
    assert _xyz(oracle, 10, (0, 3, 0)) == [3.0]                             # Length         :86-89
    assert _xyz(oracle, 11, (0, 3, 0)) == [9.0]                             # LengthSquared  :91-94
    assert _xyz(oracle, 12, (1, 2, 3), (1, 2, 3)) == [1, 4, 9]              # Mul, MulAssign :96-106
    assert _xyz(oracle, 13, (1, 2, 3), s=2.0) == [2, 4, 6]                  # MulScalar      :108-111
    assert _xyz(oracle, 14, (0, 0, 3)) == [0, 0, 1]                         # Normalize(d)   :113-132 (1/sqrt then multiply)
    assert _xyz(oracle, 15, (2, 4, 6), (1, 3, 5)) == [1, 1, 1]              # Sub, SubAssign :152-162


def test_spectrum_golden(oracle):
    # pkg/pbrt/spectrum_test.go:9-67 (Clone is Go aliasing, no counterpart: RGB is a value type here)
    sp = lambda op, a, b=(0, 0, 0), s=0.0: _kat(oracle, "spectrum_op", [op, *a, *b, s])
    assert sp(0, (1, 1, 1), (2, 2, 2)) == [3, 3, 3]        # Add, AddAssign :9-27
    assert sp(1, (1, 1, 1), s=2.0) == [3, 3, 3]            # AddScalar      :29-36
    assert sp(2, (3, 3, 3), s=3.0) == [1, 1, 1]            # DivScalar      :52-55
    assert sp(3, (3, 3, 3), (4, 4, 4)) == [12, 12, 12]     # Mul            :57-60
    assert sp(4, (1.0, 0.0, 1.0)) == [0] and sp(4, (1, 1, 1)) == [0] and sp(4, (1e-4,) * 3) == [0] and sp(4, (0, 0, 0)) == [1]  # IsBlack :62-67


def test_matches_flags_golden(oracle):
    # pkg/pbrt/reflection_test.go:9-15 with the BxDFType bits of reflection.go:286-299
    REFL, TRANS, DIFF, GLOSSY, SPEC = 1, 2, 4, 8, 16
    ALL = REFL | TRANS | DIFF | GLOSSY | SPEC
    m = lambda t, f: _kat(oracle, "matches_flags", [t, f]) == [1]
    assert m(DIFF, DIFF) and m(DIFF, DIFF | REFL) and m(REFL, DIFF | REFL) and not m(REFL, DIFF) and m(REFL, ALL)


def test_partition_primitive_info_at_golden(oracle):
    # pkg/accelerator/bvh_test.go:143-264 — the Lomuto partition behind the reference builder's SplitSAH / SplitMiddle
    # (elements are identified by centroid.x == primitiveNumber, as in the reference's table)
    part = lambda xs, start, end, pivot: [int(v) for v in _kat(oracle, "partition_at", [start, end, pivot, *xs])[:-1]]
    assert part([5, 4, 3, 2, 1], 0, 4, 2) == [1, 2, 3, 4, 5]
    assert part([5, 1, 2, 4, 3], 0, 4, 1) == [1, 3, 2, 4, 5]
