"""Independent known-answer tests for the shading rows — the rows the reference holds no
golden value for.  tests/golden/shading_kats.json is produced by tests/golden/make_shading_kats.py, a plain-Python
restatement of the cited Go lines that shares no code with oracle/ or the CUDA sources.  Here the ORACLE (CPU) and the
CUDA device functions (GPU, through the library's gopbrt_kat_eval hook) must reproduce every vector bit for bit."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from oracle_lib import OracleScene, load

HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "shading_kats.json")))
dp = C.POINTER(C.c_double)


def _f(h):
    return float.fromhex(h) if isinstance(h, str) else float(h)


def _kat_scene(gp):
    """the fixed light set of the KAT file, built through the host mirror (matrices are KAT input data)"""
    P = gp.pbrt
    lights = []
    for l in KATS["lights"]:
        if l["kind"] == "point":
            lights.append(P.Point(P.Translate(tuple(_f(x) for x in l["p"])), None, P.NewRGBSpectrum(*[_f(x) for x in l["I"]])))
        elif l["kind"] == "distant":
            lights.append(P.Distant(P.NewTransform(P.Matrix4x4()), P.NewRGBSpectrum(*[_f(x) for x in l["L"]]), tuple(_f(x) for x in l["w"])))
        else:
            xf = P.Transform(P.Matrix4x4([[_f(x) for x in row] for row in l["m"]]), P.Matrix4x4([[_f(x) for x in row] for row in l["minv"]]))
            if l["shape"] == "sphere":
                shape = P.NewSphereShape("kat", xf, False, _f(l["radius"]))
            else:
                shape = P.Disk(xf, _f(l["height"]), _f(l["radius"]), _f(l["inner"]), 360.0)
            lights.append(P.DiffuseAreaLight(P.NewTransform(P.Matrix4x4()), None, P.NewRGBSpectrum(*[_f(x) for x in l["L"]]), 1, shape, bool(l["two_sided"])))
    far = P.GeometricPrimitive(P.NewSphereShape("far", P.Translate((1000.0, 1000.0, 1000.0)), False, 1.0),
                               P.MatteMaterial(P.ConstantSpectrumTexture(P.NewSpectrum(0.5)), P.ConstantFloatTexture(0.0)))
    return P.NewScene(P.NewBVH([far], 4, P.SplitSAH), lights)


def _check(case, got, n):
    want = np.array([_f(x) for x in case["out"]])
    assert n == len(want), f"{case['fn']} {case['cite']}: {n} values, expected {len(want)}"
    got = np.array(got[:n])
    same = (got == want) | (np.isnan(got) & np.isnan(want))
    assert same.all(), (f"{case['fn']} ({case['cite']}) in={[_f(x) for x in case['in']]}\n got  {[float(x).hex() for x in got]}\n"
                        f" want {[float(x).hex() for x in want]}")


def test_kat_file_is_current():
    # the committed vectors are what the generator produces today (the generator is the specification)
    import importlib.util
    import tempfile
    spec = importlib.util.spec_from_file_location("make_shading_kats", os.path.join(HERE, "golden", "make_shading_kats.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    old = mod.HERE
    with tempfile.TemporaryDirectory() as td:
        mod.HERE = td
        mod.main()
        fresh = json.load(open(os.path.join(td, "shading_kats.json")))
    mod.HERE = old
    assert fresh == KATS
    assert len(KATS["cases"]) > 250 and {c["fn"] for c in KATS["cases"]} >= {
        "fr_dielectric", "oren_nayar_f", "fresnel_specular_sample_f", "light_sample_li", "sample_discrete_uniform", "film_add_sample",
        "stratified_start_pixel", "camera_ray", "rng_u32", "lambert_sample_f"}


def test_oracle_reproduces_independent_kats(gp):
    lib = load()
    lib.oracle_kat_eval.argtypes = [C.c_void_p, C.c_char_p, dp, C.c_int, dp, C.c_int]
    lib.oracle_kat_eval.restype = C.c_int
    o = OracleScene(_kat_scene(gp), 1)
    out = (C.c_double * 512)()
    for case in KATS["cases"]:
        vin = (C.c_double * len(case["in"]))(*[_f(x) for x in case["in"]])
        n = lib.oracle_kat_eval(o.h, case["fn"].encode(), vin, len(case["in"]), out, 512)
        assert n >= 0, f"oracle has no KAT entry for {case['fn']} with {len(case['in'])} inputs"
        _check(case, list(out), n)
    o.close()


@pytest.mark.gpu
def test_cuda_device_functions_reproduce_independent_kats(gp, dev):
    g = gp.pbrt.GpuScene(dev, _kat_scene(gp))
    lib = dev.lib
    out = (C.c_double * 512)()
    seen = set()
    for case in KATS["cases"]:
        fid = gp.abi.KAT_IDS.get(case["fn"])
        assert fid is not None, case["fn"]
        vin = (C.c_double * len(case["in"]))(*[_f(x) for x in case["in"]])
        n = lib.gopbrt_kat_eval(g.h, fid, vin, len(case["in"]), out, 512)
        assert n >= 0, f"gopbrt_kat_eval({case['fn']}) rc={n}: {dev.error()}"
        _check(case, list(out), n)
        seen.add(case["fn"])
    assert seen == set(gp.abi.KAT_IDS)
    g.close()
