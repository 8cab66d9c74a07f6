"""The DirectLighting integrator, composed, in plain Python floats — the companion of make_path_golden.py (same rules: written
from the Go source, nothing from oracle/ or go-pbrt_b200/csrc/; the per-function pieces come from make_shading_kats.py, the
scene / shapes / samplers / film from make_path_golden.py).

What is new here (directlighting.go:62-119, integrator.go:21-58,352-434, glass.go:38-73, reflection.go:404-456,538-575):
 * DirectLighting.Li calls ComputeScatteringFunctions with allowMultipleLobes = FALSE: smooth glass is then TWO lobes — a
   SpecularReflection with a dielectric Fresnel term, typed Reflection|Diffuse, and a SpecularTransmission typed
   Transmission|Specular whose SampleF answers sampledType 0 — and BSDF.SampleF picks among the lobes that match the flags;
 * UniformSampleAllLights (every light once: the per-tile sampler clones hold no sample arrays, sampler.go:62-69, pixel.go:34-42)
   and UniformSampleOneLight with a nil distribution (lightNum = min(u * n, n - 1));
 * the specular recursion: SpecularReflect asks for Reflection|Specular, which NO lobe of the three materials is a subset of (the
   mirror's lobe is typed Diffuse) — it draws its 2-D sample and returns black; SpecularTransmit follows the glass's transmission
   lobe along the LOCAL wi; `depth` grows by two per level (Li passes depth+1, the helper passes depth+1 again).

    python tests/golden/make_direct_golden.py        # rewrites tests/golden/direct_golden.json
"""
import importlib
import importlib.util
import json
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
_spec = importlib.util.spec_from_file_location("make_path_golden", os.path.join(HERE, "make_path_golden.py"))
M = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(M)
K = M.K
Z3 = M.Z3
REFL, TRANS, DIFF, SPEC, ALL, NON_SPECULAR = M.REFL, M.TRANS, M.DIFF, M.SPEC, M.ALL, M.NON_SPECULAR

UNIFORM_SAMPLE_ALL, UNIFORM_SAMPLE_ONE = 1, 2   # directlighting.go:12-15
MAX_DEPTH = 5
TILE = 5
CASES = {"all": UNIFORM_SAMPLE_ALL, "one": UNIFORM_SAMPLE_ONE}


def scene_and_integrator(gp, strategy):
    """the scene, camera and sampler of make_path_golden.py under integrator.NewDirectLighting(strategy, 5, ...)"""
    scene, path = M.scene_and_integrator(gp)
    return scene, gp.pbrt.NewDirectLighting(strategy, MAX_DEPTH, path.GetCamera(), path.GetSampler(), None)


# ---------------------------------------------------------------- reflection.go:128-277 with any number of lobes
class BSDF:
    def __init__(self, hit):  # NewBSDF + the materials with allowMultipleLobes = false
        self.ns, self.ng = hit["ns"], hit["n"]
        self.ss = K.v_normalized(hit["sh_dpdu"])
        self.ts = K.v_cross(self.ns, self.ss)
        m = hit["disk"]["mat"]
        self.lobes = []
        if m["kind"] == "matte":
            if not M.black(m["kd"]):
                if m.get("sigma", 0.0) == 0:
                    self.lobes.append(dict(kind="lambert", type=REFL | DIFF, r=m["kd"]))
                else:   # OrenNayar (reflection.go:616-652)
                    self.lobes.append(dict(kind="orennayar", type=REFL | DIFF, r=m["kd"], sigma=m["sigma"]))
        elif m["kind"] == "mirror":
            if not M.black(m["kr"]):
                self.lobes.append(dict(kind="specrefl", type=REFL | DIFF, r=m["kr"], fresnel=None))
        else:  # glass.go:38-73: R or T non-black, uRough == vRough == 0, allowMultipleLobes false
            if not (M.black(m["R"]) and M.black(m["T"])):
                if not M.black(m["R"]):
                    self.lobes.append(dict(kind="specrefl", type=REFL | DIFF, r=m["R"], fresnel=(1.0, m["eta"])))
                if not M.black(m["T"]):
                    self.lobes.append(dict(kind="spectrans", type=TRANS | SPEC, t=m["T"], etaA=1.0, etaB=m["eta"]))

    def to_local(self, v):
        return [K.v_dot(v, self.ss), K.v_dot(v, self.ts), K.v_dot(v, self.ns)]

    def to_world(self, v):  # :151-157
        return [self.ss[i] * v[0] + self.ts[i] * v[1] + self.ns[i] * v[2] for i in range(3)]

    def matching(self, flags):
        return [lb for lb in self.lobes if (lb["type"] & flags) == lb["type"]]

    @staticmethod
    def lobe_f(lb, wo, wi):
        if lb["kind"] == "lambert":
            return [c * K.INV_PI for c in lb["r"]]
        if lb["kind"] == "orennayar":
            return K.oren_nayar(lb["r"], lb["sigma"], wo, wi)
        return list(Z3)

    @staticmethod
    def lobe_pdf(lb, wo, wi):
        if lb["kind"] in ("lambert", "orennayar"):
            return abs(wi[2]) * K.INV_PI if wo[2] * wi[2] > 0 else 0.0
        return 0.0

    @staticmethod
    def lobe_sample_f(lb, wo, u):
        if lb["kind"] in ("lambert", "orennayar"):   # sampleF :305-314
            wi = K.cosine_sample_hemisphere(u, M.COS, M.SIN)
            if wo[2] < 0:
                wi[2] *= -1
            return BSDF.lobe_f(lb, wo, wi), wi, BSDF.lobe_pdf(lb, wo, wi), 0
        if lb["kind"] == "specrefl":  # :557-562; FresnelNoOp :383-385 / FresnelDielectric :400-402
            wi = [-wo[0], -wo[1], wo[2]]
            fr = 1.0 if lb["fresnel"] is None else K.fr_dielectric(wi[2], lb["fresnel"][0], lb["fresnel"][1])
            return [(fr * c) / abs(wi[2]) for c in lb["r"]], wi, 1.0, 0
        # SpecularTransmission.SampleF :426-446 (mode == Radiance)
        entering = wo[2] > 0
        etaI, etaT = (lb["etaA"], lb["etaB"]) if entering else (lb["etaB"], lb["etaA"])
        wi = K.refract(wo, K.face_forward([0.0, 0.0, 1.0], wo), etaI / etaT)
        if wi is None:
            return list(Z3), list(Z3), 0.0, 0
        fr = K.fr_dielectric(wi[2], lb["etaA"], lb["etaB"])
        ft = [c * (1 - fr) for c in lb["t"]]
        s = (etaI * etaI) / (etaT * etaT)
        ft = [c * s for c in ft]
        return [c / abs(wi[2]) for c in ft], wi, 1.0, 0

    def f(self, wo_w, wi_w, flags):  # :170-187
        wi, wo = self.to_local(wi_w), self.to_local(wo_w)
        if wo[2] == 0.0:
            return list(Z3)
        reflect = K.v_dot(wi_w, self.ng) * K.v_dot(wo_w, self.ng) > 0
        f = list(Z3)
        for lb in self.matching(flags):
            if (reflect and lb["type"] & REFL > 0) or (not reflect and lb["type"] & TRANS > 0):
                lf = self.lobe_f(lb, wo, wi)
                f = [f[i] + lf[i] for i in range(3)]
        return f

    def pdf(self, wo_w, wi_w, flags):  # :255-277
        if not self.lobes:
            return 0.0
        wo, wi = self.to_local(wo_w), self.to_local(wi_w)
        if wo[2] == 0:
            return 0.0
        pdf, n = 0.0, 0
        for lb in self.matching(flags):
            n += 1
            pdf += self.lobe_pdf(lb, wo, wi)
        return 0.0 if n <= 0 else pdf / float(n)

    def sample_f(self, wo_w, u, flags):  # :189-253; the LOCAL wi is what it returns
        none = (list(Z3), list(Z3), 0.0, 0)
        cand = self.matching(flags)
        n = len(cand)
        if n == 0:
            return none
        comp = K.go_min(math.floor(u[0] * float(n)), float(n) - 1)
        lb = cand[int(comp)]
        ur = [K.go_min(u[0] * float(n) - comp, K.ONE_MINUS_EPSILON), u[1]]
        wo = self.to_local(wo_w)
        if wo[2] == 0.0:
            return none
        f, wi, pdf, st = self.lobe_sample_f(lb, wo, ur)
        if pdf == 0.0:
            return none
        wi_w = self.to_world(wi)
        if lb["type"] & SPEC <= 0 and n > 1:
            for other in cand:
                if other is not lb:
                    pdf += self.lobe_pdf(other, wo, wi)
        if n > 1:
            pdf /= float(n)
        if lb["type"] & SPEC == 0 and n > 1:
            reflect = K.v_dot(wi_w, self.ng) * K.v_dot(wo_w, self.ng) > 0
            f = list(Z3)
            for other in cand:
                if (reflect and other["type"] & REFL > 0) or (not reflect and other["type"] & TRANS > 0):
                    lf = self.lobe_f(other, wo, wi)
                    f = [f[i] + lf[i] for i in range(3)]
        return f, wi, pdf, st


# ---------------------------------------------------------------- integrator.go:21-76, directlighting.go:62-119, integrator.go:352-434
def sample_lights(hit, bsdf, strategy, sc, scene, smp, stats):
    n = len(sc["lights"])
    L = list(Z3)
    if strategy == UNIFORM_SAMPLE_ALL:   # Get2DArray answers nil on a clone: one sample per light, no division
        for light in sc["lights"]:
            u_light = smp.get2d()
            u_scattering = smp.get2d()
            Ld = M.estimate_direct(hit, bsdf, light, u_light, u_scattering, scene, stats)
            L = [L[i] + Ld[i] for i in range(3)]
        return L
    num = int(K.go_min(smp.get1d() * float(n), float(n - 1)))   # lightDistrib == nil (directlighting.go:97)
    u_light = smp.get2d()
    u_scattering = smp.get2d()
    Ld = M.estimate_direct(hit, bsdf, sc["lights"][num], u_light, u_scattering, scene, stats)
    stats["max_direct"] = max(stats["max_direct"], max(Ld))   # DivScalar(lightPdf) dropped; > 10 panics (integrator.go:71-74)
    return Ld


def direct_li(o, w, sc, scene, smp, stats, depth=0):
    L = list(Z3)
    hit = scene.intersect(o, w, M.INF)
    if hit is None:
        return L   # every Light.Le is a zero spectrum (light.go:31-33)
    bsdf = BSDF(hit)
    L = [L[i] + 0.0 for i in range(3)]   # si.Le(si.Wo): no primitive carries an area light
    if sc["lights"]:
        Ld = sample_lights(hit, bsdf, sc["strategy"], sc, scene, smp, stats)
        L = [L[i] + Ld[i] for i in range(3)]
    if depth + 1 < sc["max_depth"]:
        d = depth + 1
        # SamplerIntegratorSpecularReflect (integrator.go:352-384): the sample is drawn, no lobe matches Reflection|Specular
        f, wi, pdf, _ = bsdf.sample_f(hit["wo"], smp.get2d(), REFL | SPEC)
        assert pdf == 0.0 and M.black(f)
        stats["spec_reflect_calls"] += 1
        L = [L[i] + 0.0 for i in range(3)]
        # SamplerIntegratorSpecularTransmit (integrator.go:386-434)
        f, wi, pdf, _ = bsdf.sample_f(hit["wo"], smp.get2d(), TRANS | SPEC)
        Lt = list(Z3)
        if pdf > 0 and not M.black(f) and abs(K.v_dot(wi, hit["ns"])) != 0.0:
            stats["spec_transmit_rays"] += 1
            ro = K.offset_ray_origin(hit["p"], hit["perr"], hit["n"], wi)   # si.SpawnRay(wi), wi LOCAL
            Li = direct_li(ro, wi, sc, scene, smp, stats, d + 1)
            k = abs(K.v_dot(wi, hit["ns"])) / pdf
            Lt = [(f[i] * Li[i]) * k for i in range(3)]
            stats["max_level"] = max(stats["max_level"], (d + 1) // 2)
        L = [L[i] + Lt[i] for i in range(3)]
    return L


def render(sc, tile, strategy):
    sc = dict(sc, strategy=strategy, max_depth=MAX_DEPTH)
    extra = dict(spec_reflect_calls=0, spec_transmit_rays=0, max_level=0)

    def li(o, w, sc_, scene, smp, stats):
        for k, v in extra.items():
            stats.setdefault(k, v)
        return direct_li(o, w, sc_, scene, smp, stats)

    saved = M.path_li
    M.path_li = li   # renderWorker (integrator.go:228-289) is integrator-agnostic: s.Li
    try:
        return M.render(sc, tile)
    finally:
        M.path_li = saved


def main():
    gp = importlib.import_module("go-pbrt_b200")
    out = dict(note="made by tests/golden/make_direct_golden.py (plain-Python restatement of the composed DirectLighting hot path); "
                    "film = [y][x][X, Y, Z, filterWeightSum] as float.hex()", cases={})
    for name, strategy in CASES.items():
        scene, integ = scene_and_integrator(gp, strategy)
        sc = M.plain_scene(scene, M.scene_and_integrator(gp)[1])
        film, st = render(sc, TILE, strategy)
        assert st["max_direct"] <= 10.0
        print(f"{name}: camera {st['camera']}, closest {st['closest']}, shadow {st['shadow']}, area-light estimates {st['nondelta']}, "
              f"transmitted rays {st['spec_transmit_rays']} (deepest level {st['max_level']}), SpecularReflect calls {st['spec_reflect_calls']}")
        out["cases"][name] = dict(strategy=strategy, tile=TILE, max_depth=MAX_DEPTH, rays=[st["camera"], st["closest"], st["shadow"]],
                                  nondelta_estimates=st["nondelta"],
                                  coverage=dict(transmitted_rays=st["spec_transmit_rays"], deepest_level=st["max_level"]),
                                  film=[[[v.hex() for v in p] for p in row] for row in film])
    with open(os.path.join(HERE, "direct_golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote direct_golden.json")


if __name__ == "__main__":
    main()
