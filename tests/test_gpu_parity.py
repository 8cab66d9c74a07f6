"""Parity of the CUDA path (through the C ABI) against the oracle.  Run with -m gpu on a B200.
Bar (BASELINE.json north_star): hit/miss and primitive ids bit-exact, t to 1e-5 relative (we assert bit-exact),
film RMSE below a stated threshold (we assert bit-exact in STRICT mode; tolerances are written where they apply)."""
import numpy as np
import pytest

from oracle_lib import OracleScene, camera_rays

pytestmark = pytest.mark.gpu


def _cmp_closest(g, o, what, reference_order=False):
    """hit/miss, t, primitive id, hit point and normal bit-exact.  Against the parity-spec oracle (accel 1 / 2) nothing is
    tolerated: the closest hit is defined order-independently.  reference_order=True (the oracle running the
    reference's own tree, visit order and running tMax, accel 0): two primitives hit at the very same t go to whichever the
    reference's tree visits first, so up to n/20000 primitive ids may differ; the count is printed."""
    gp_, gt, gpnt, gn = g
    op, ot, opnt, on = o
    assert np.array_equal(gp_ >= 0, op >= 0), f"{what}: hit/miss differs on {np.count_nonzero((gp_ >= 0) != (op >= 0))} rays"
    assert np.array_equal(gt, ot), f"{what}: t differs on {np.count_nonzero(gt != ot)} rays"
    ties = gp_ != op
    print(f"[parity] {what}: {len(op)} rays, {np.count_nonzero(op >= 0)} hits, {np.count_nonzero(ties)} primitive-id ties")
    allowed = max(1, len(op) // 20000) if reference_order else 0
    assert np.count_nonzero(ties) <= allowed, f"{what}: {np.count_nonzero(ties)} primitive ids differ of {len(op)}"
    same = ~ties
    assert np.array_equal(gpnt[same], opnt[same]) and np.array_equal(gn[same], on[same]), f"{what}: hit point / normal differ"


TEST_RAYS = [((0, 0, 0), (0, 0, 1.0)), ((0, 0, 0), (0, 0, -1.0)), ((0, 0, 500), (0, 0, -1.0)), ((10, 10, 500), (0, 0, -1.0))]


def test_reference_bvh_test_vectors(gp, dev):
    # pkg/accelerator/bvh_test.go:43-141, simple_test.go:40-108 replayed on the GPU aggregate
    P = gp.pbrt
    for max_prims in (255, 2):
        scene = P.NewScene(P.NewBVH(gp.scenes.bvh_test_primitives(), max_prims, P.SplitSAH), [])
        g = P.GpuScene(dev, scene)
        o = [r[0] for r in TEST_RAYS]
        d = [r[1] for r in TEST_RAYS]
        prim, t, p, n = g.Intersect(o, d)
        assert list(prim) == [0, -1, 1, 2]
        assert list(t[[0, 2, 3]]) == [4.0, 489.0, 489.0]
        assert list(g.IntersectP(o, d)) == [True, False, True, True]
        inv = 1.0 / np.sqrt(3.0)
        prim, t, p, n = g.Intersect([(15, 15, 15)], [(-inv, -inv, -inv)])
        assert prim[0] == 2 and list(p[0]) == [10 + inv] * 3 and t[0] == 7.6602540378443855
        g.close()


def test_empty_and_zero_rays(gp, dev):
    P = gp.pbrt
    g = P.GpuScene(dev, P.NewScene(P.NewBVH([], 4, P.SplitSAH), []))
    prim, t, p, n = g.Intersect([(0, 0, 0)], [(0, 0, 1)])
    assert prim[0] == -1 and np.isinf(t[0])
    assert not g.IntersectP([(0, 0, 0)], [(0, 0, 1)])[0]
    prim, t, p, n = g.Intersect(np.zeros((0, 3)), np.zeros((0, 3)))
    assert len(prim) == 0
    g.close()


def test_config1_primary_rays_bit_exact(gp, dev):
    # every primary ray of the README scene at 1920x1080 on a 3-pixel lattice + all pixels of a 480x270 render,
    # against BOTH the reference-BVH oracle (accel 0) and the parity-spec oracle (accel 1)
    scene, integ = gp.scenes.config1()
    g = gp.pbrt.GpuScene(dev, scene)
    ys, xs = np.mgrid[0:1080:3, 0:1920:3]
    o, d = camera_rays(integ, xs.ravel(), ys.ravel())
    res = g.Intersect(o, d)
    for accel in (0, 1):
        s = OracleScene(scene, accel)
        _cmp_closest(res, s.intersect(o, d), f"config1 primary accel={accel}", reference_order=accel == 0)
        assert np.array_equal(g.IntersectP(o, d), s.intersect_p(o, d))
        s.close()
    assert np.count_nonzero(res[0] >= 0) > 0.5 * len(o)
    g.close()


def _secondary_rays(oscene, o, d, seed):
    """shadow-like and bounce-like rays leaving the first hit points (origins ON the surfaces)"""
    prim, t, p, n = oscene.intersect(o, d)
    m = prim >= 0
    p = p[m]
    rng = np.random.default_rng(seed)
    dirs = rng.normal(size=p.shape)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    return p, dirs


def test_config1_secondary_rays_bit_exact(gp, dev):
    scene, integ = gp.scenes.config1()
    g = gp.pbrt.GpuScene(dev, scene)
    s = OracleScene(scene, 1)
    s0 = OracleScene(scene, 0)
    ys, xs = np.mgrid[0:1080:7, 0:1920:7]
    o, d = camera_rays(integ, xs.ravel(), ys.ravel())
    p, dirs = _secondary_rays(s, o, d, 1)
    _cmp_closest(g.Intersect(p, dirs), s.intersect(p, dirs), "config1 secondary")
    _cmp_closest(g.Intersect(p, dirs), s0.intersect(p, dirs), "config1 secondary vs reference BVH", reference_order=True)
    to_light = np.array([50.0, 20.0, 50.0]) - p  # shadow rays toward the point light, tMax = 1 - ShadowEpsilon
    assert np.array_equal(g.IntersectP(p, to_light, 0.9999), s.intersect_p(p, to_light, 0.9999))
    assert np.array_equal(g.IntersectP(p, to_light, 0.9999), s0.intersect_p(p, to_light, 0.9999))
    g.close(); s.close(); s0.close()


def test_mixed_scene_rays_bit_exact(gp, dev):
    # spheres (plain, reversed, TransformedPrimitive with rotation), disk, triangles; 200k random + camera + secondary rays
    scene = gp.scenes.mixed_test_scene(300)
    g = gp.pbrt.GpuScene(dev, scene)
    s = OracleScene(scene, 1)
    rng = np.random.default_rng(3)
    n = 200000
    o = rng.uniform(-15, 15, size=(n, 3)); o[:, 1] = rng.uniform(0, 12, size=n)
    d = rng.normal(size=(n, 3))
    tm = np.where(rng.uniform(size=n) < 0.3, rng.uniform(1, 30, size=n), np.inf)
    _cmp_closest(g.Intersect(o, d, tm), s.intersect(o, d, tm), "mixed random")
    assert np.array_equal(g.IntersectP(o, d, tm), s.intersect_p(o, d, tm))
    p, dirs = _secondary_rays(s, o, d, 5)
    _cmp_closest(g.Intersect(p, dirs), s.intersect(p, dirs), "mixed secondary")
    # axis-parallel rays (infinite invDir components)
    ax = np.eye(3)[rng.integers(0, 3, size=5000)] * rng.choice([-1.0, 1.0], size=(5000, 1))
    _cmp_closest(g.Intersect(o[:5000], ax), s.intersect(o[:5000], ax), "mixed axis-parallel")
    g.close(); s.close()


def _render_both(gp, dev, scene, integ, tile, accel=1, **kw):
    g = gp.pbrt.GpuScene(dev, scene)
    st = gp.pbrt.Render(g, integ, tile, **kw)
    film = integ.GetCamera().GetFilm().pixels.copy()
    s = OracleScene(scene, accel)
    ofilm, ost = s.render(integ, tile, mode=kw.get("mode", 0))  # one lane per tile: FAST callers pass groups=1 for bit parity
    assert st["efloat_panics"] == 0 and st["stack_overflows"] == 0
    g.close(); s.close()
    return film, st, ofilm, ost


def _assert_film_equal(film, ofilm, st, ost, what, exact=True):
    """STRICT-mode film parity.  exact=True: every pixel bit-identical and identical ray counts.  exact=False (scenes
    where two tree topologies may break an exact t tie differently): the SURVEY §8d tier-2 bar — filterWeightSum exact,
    at most 0.01 % of pixels (>= 1) may hold an outlier from a flipped path, all others bit-identical."""
    assert st["camera_rays"] == ost["camera_rays"]
    assert np.array_equal(film[..., 3], ofilm[..., 3]), f"{what}: filterWeightSum differs"
    bad = np.count_nonzero(np.any(film != ofilm, axis=2))
    npx = film.shape[0] * film.shape[1]
    if exact:
        for k in ("closest_rays", "shadow_rays"):
            assert st[k] == ost[k], f"{what}: {k} {st[k]} != {ost[k]}"
        assert bad == 0, f"{what}: {bad} of {npx} pixels differ"
    else:
        assert bad <= max(1, npx // 10000), f"{what}: {bad} of {npx} pixels differ"
        for k in ("closest_rays", "shadow_rays"):
            assert abs(st[k] - ost[k]) <= max(20, ost[k] // 10000), f"{what}: {k} {st[k]} vs {ost[k]}"


@pytest.mark.parametrize("tile", [16, 1])
def test_config1_film_bit_exact(gp, dev, tile):
    # README scene, Stratified(4,4), Path(maxDepth 10): STRICT mode reproduces pbrt.Render(…, tileSize) sample for sample.
    # Tier-2 bar of SURVEY §8d is relative RMSE <= 1e-6; we hold the stricter bit-exact bar, vs the reference-BVH oracle.
    scene, integ = gp.scenes.config1(W=320, H=180)
    film, st, ofilm, ost = _render_both(gp, dev, scene, integ, tile, accel=0)
    _assert_film_equal(film, ofilm, st, ost, f"config1 tile={tile}")
    assert st["camera_rays"] == 320 * 180 * 15


@pytest.mark.parametrize("tile", [8, 1])
def test_mixed_scene_film_bit_exact(gp, dev, tile):
    scene = gp.scenes.mixed_test_scene(150)
    integ = gp.scenes.test_integrator(160, 96, spp=(3, 3), maxDepth=8)
    film, st, ofilm, ost = _render_both(gp, dev, scene, integ, tile)
    _assert_film_equal(film, ofilm, st, ost, f"mixed tile={tile}")


def test_jittered_and_ragged_tiles_film_bit_exact(gp, dev):
    # jitter=True consumes RNG in StartPixel; 75x41 is not a multiple of the tile size (ragged last tiles)
    scene = gp.scenes.mixed_test_scene(60, seed=11)
    integ = gp.scenes.test_integrator(75, 41, spp=(2, 3), maxDepth=5, jitter=True, ndims=3)
    film, st, ofilm, ost = _render_both(gp, dev, scene, integ, 16)
    _assert_film_equal(film, ofilm, st, ost, "jittered ragged")


def test_config2_cornell_film_bit_exact(gp, dev):
    scene, integ = gp.scenes.config2(W=160, H=90, spp=(3, 3))
    film, st, ofilm, ost = _render_both(gp, dev, scene, integ, 1)
    _assert_film_equal(film, ofilm, st, ost, "config2")
    assert st["radiance_gt10"] == ost["radiance_gt10"]


def test_flat_and_bvh_aggregates_agree_bitwise(gp, dev, monkeypatch):
    # scenes of <= 64 primitives are answered by the flat table (k_trace_flat), larger ones by the BVH (k_trace).  The
    # closest hit is defined order-independently, so both give the same film bit for bit — on config 2,
    # whose boxes rest on the floor (coplanar faces: the reference's running-tMax rule depends on the visit order there)
    scene, integ = gp.scenes.config2(W=160, H=90, spp=(3, 3))
    out = []
    for no_flat in (False, True):
        if no_flat:
            monkeypatch.setenv("GOPBRT_NO_FLAT", "1")
        g = gp.pbrt.GpuScene(dev, scene)
        st = gp.pbrt.Render(g, integ, 1, flags=gp.abi.FLAG_COUNT_TRAVERSAL)
        out.append((integ.GetCamera().GetFilm().pixels.copy(), st))
        g.close()
    (f0, s0), (f1, s1) = out
    assert np.array_equal(f0, f1)
    for k in ("closest_rays", "shadow_rays", "dead_mis_rays", "radiance_gt10"):
        assert s0[k] == s1[k], k
    assert s0["nodes_visited"] != s1["nodes_visited"]  # two different aggregates did run


def test_last_sample_in_place_equals_retirement_through_raygen(gp, dev, monkeypatch):
    # FAST mode + uniform footprint + Path: a lane's last sample is not retired through the regeneration queue — the film
    # fold adds pad + L (k_group_sums<LAST>, or k_fold_last when a pass does not hold whole tiles).  Same additions in the
    # same order as the retirement inside raygen (GOPBRT_NO_LAST_IN_PLACE=1): bit-identical films and identical counters,
    # with one lane per pixel, several lanes per pixel (multi-sample lanes), one sample per lane, and split passes.
    scene, integ = gp.scenes.config2(W=96, H=54, spp=(4, 4))
    g = gp.pbrt.GpuScene(dev, scene)
    for kw in (dict(groups=1), dict(groups=4), dict(groups=15), dict(groups=0), dict(groups=4, max_lanes=7001), dict(groups=3, rank=1, world=2)):
        films, stats = [], []
        for off in (False, True):
            if off:
                monkeypatch.setenv("GOPBRT_NO_LAST_IN_PLACE", "1")
            else:
                monkeypatch.delenv("GOPBRT_NO_LAST_IN_PLACE", raising=False)
            st = gp.pbrt.Render(g, integ, 1, mode=gp.abi.MODE_FAST, **kw)
            films.append(integ.GetCamera().GetFilm().pixels.copy())
            stats.append(st)
        assert np.array_equal(films[0], films[1]), kw
        for k in ("camera_rays", "closest_rays", "shadow_rays", "nan_samples", "shaded_lanes"):
            assert stats[0][k] == stats[1][k], (kw, k)
    monkeypatch.delenv("GOPBRT_NO_LAST_IN_PLACE", raising=False)
    g.close()


def test_fast_mode_film_bit_exact_and_statistically_close_to_strict(gp, dev):
    scene = gp.scenes.mixed_test_scene(80, seed=5)
    integ = gp.scenes.test_integrator(96, 64, spp=(4, 4), maxDepth=5)
    film, st, ofilm, ost = _render_both(gp, dev, scene, integ, 1, mode=gp.abi.MODE_FAST, groups=1)
    _assert_film_equal(film, ofilm, st, ost, "fast mode")


@pytest.mark.parametrize("depth", ["direct", "full"])
@pytest.mark.parametrize("config", ["config1", "config2"])
def test_fast_mode_is_statistically_equivalent_to_strict(gp, dev, config, depth):
    """SURVEY §8d tier 3.  FAST (counter-based streams per (pixel, sample): the mode that splits by sample index) and STRICT
    (the unmodified reference's per-tile streams) draw different random numbers for the same estimator, so their films
    differ by Monte-Carlo noise only.  The per-sample films X_s (FAST, rank s of world = spp renders exactly sample s)
    give the per-pixel sample standard deviation sigma; with n = spp - 1 samples per pixel, D = (FAST - STRICT) / n:
      * image RMSE of D             <= 3 * rms(sigma) / sqrt(n)                    (the survey's bar)
      * |D| > 4 * sigma * sqrt(2/n)  on at most 1 % of the pixels with sigma > 0   (two independent n-sample means)
      * |mean(FAST) - mean(STRICT)| <= max(0.1 % of the mean, 4 standard errors)  (0.1 % is the survey's bar at 1080p;
        at this test's 160 x 90 the standard error of the image mean is larger than that and is what binds)
    and the weights (filterWeightSum) are identical.
    depth = "direct": maxDepth 2, i.e. the direct light at the first hit — a bounded integrand, every pixel counts.
    depth = "full":   the configuration's own maxDepth 10.  The reference's estimator is heavy-tailed there: BSDF.SampleF
        hands back the LOCAL wi, so beta *= f * |wi . ns| / pdf keeps the 1 / |cos theta| of every specular
        lobe (mirror, glass) — single samples of 1e8 .. 1e14 occur in both modes and no sample mean converges, so the
        three bars above are not defined.  The test then uses statistics that do not need moments: the pixels whose two
        means both stay below 10 (the reference's own sanity bound, integrator.go:73-75; the rule is symmetric in the
        two films, so it cannot favour one) must be most of the image, their mean difference must vanish within four
        EMPIRICAL standard errors (std of D over those pixels, 4 pixels per independent footprint), and the median of
        |D| must match what sigma predicts for two independent n-sample means (0.6745 * sigma * sqrt(2/n)) within a
        factor of two.
    """
    P = gp.pbrt
    scene, integ = getattr(gp.scenes, config)(W=160, H=90)
    if depth == "direct":
        integ = P.NewPath(2, integ.GetCamera(), integ.GetSampler(), None, 1, P.Uniform)
    spp = integ.GetSampler().GetSamplesPerPixel()
    n = spp - 1
    g = P.GpuScene(dev, scene)
    P.Render(g, integ, 1)
    strict = integ.GetCamera().GetFilm().pixels.copy()
    P.Render(g, integ, 1, mode=gp.abi.MODE_FAST)
    fast = integ.GetCamera().GetFilm().pixels.copy()
    s1 = np.zeros(strict.shape[:2] + (3,))
    s2 = np.zeros_like(s1)
    for s in range(1, spp):  # rank s of world spp owns exactly sample index s (rank 0 owns none: indices start at 1)
        P.Render(g, integ, 1, mode=gp.abi.MODE_FAST, rank=s, world=spp)
        x = integ.GetCamera().GetFilm().pixels[..., :3]
        s1 += x
        s2 += x * x
    g.close()
    assert np.array_equal(fast[..., 3], strict[..., 3])
    assert np.allclose(s1, fast[..., :3], rtol=1e-12, atol=1e-300)  # the per-sample films are the FAST film, split
    F, S = fast[..., :3] / n, strict[..., :3] / n
    var = np.maximum(s2 / n - (s1 / n) ** 2, 0.0) * n / (n - 1)
    sigma = np.sqrt(var)
    if depth == "full":
        keep = np.broadcast_to(((F <= 10.0) & (S <= 10.0)).all(axis=2, keepdims=True), F.shape)
        D = (F - S)[keep]
        se = D.std() / np.sqrt(D.size / 4.0)
        lit = sigma[keep] > 0
        mad = np.median(np.abs(D[lit]))
        pred = np.median(0.6745 * sigma[keep][lit] * np.sqrt(2.0 / n))
        print(f"[tier3] {config} full: n={n} pixels kept {100 * keep.mean():.1f}% mean FAST={F[keep].mean():.6g} STRICT={S[keep].mean():.6g} "
              f"diff={D.mean():.4g} 4SE={4 * se:.4g} median|D|={mad:.4g} predicted={pred:.4g}")
        # how many pixels stay below the bound is itself a statistic both modes must agree on (config 2, a closed room with a
        # mirror and a glass sphere, keeps under half of them at 63 spp)
        kf, ks = (F <= 10.0).all(axis=2).mean(), (S <= 10.0).all(axis=2).mean()
        assert keep.mean() >= 0.25, f"only {100 * keep.mean():.1f}% of the pixels are free of heavy-tail samples"
        npix = F.shape[0] * F.shape[1]
        assert abs(kf - ks) <= 4.0 * np.sqrt(2.0 * max(kf * (1 - kf), 1e-4) / npix), (kf, ks)  # two independent binomial fractions
        assert abs(D.mean()) <= 4.0 * se
        assert 0.5 * pred <= mad <= 2.0 * pred
        return
    D = F - S
    rmse = np.sqrt(np.mean(D ** 2))
    bar = 3.0 * np.sqrt(np.mean(var)) / np.sqrt(n)
    lit = sigma > 0
    out = np.count_nonzero(np.abs(D[lit]) > 4.0 * sigma[lit] * np.sqrt(2.0 / n)) / max(1, np.count_nonzero(lit))
    mean_f, mean_s = F.mean(), S.mean()
    se = np.sqrt(2.0 * np.mean(var) / n / (D.size / 4.0))  # standard error of the difference of the two image means
    print(f"[tier3] {config} direct: n={n} RMSE(D)={rmse:.4g} bar={bar:.4g} outliers={100 * out:.2f}% "
          f"mean FAST={mean_f:.6g} STRICT={mean_s:.6g} diff={abs(mean_f - mean_s) / mean_s * 100:.3f}% 4SE={4 * se / mean_s * 100:.3f}%")
    assert rmse <= bar
    assert out <= 0.01
    assert abs(mean_f - mean_s) <= max(1e-3 * mean_s, 4.0 * se)


def test_path_power_light_strategy_and_spatial(gp, dev):
    # LightSampleStrategy Power (lightdistribution.go:44-68), bugs included: the distribution is 2n zeros, SampleDiscrete's
    # pdf is 0, so UniformSampleOneLight returns black after its one Get1D — no shadow ray is ever traced and the film
    # carries weights only.  Spatial has no distribution at all in the reference (nil): GOPBRT_ERR_UNSUPPORTED.
    P = gp.pbrt
    scene = gp.scenes.mixed_test_scene(80, seed=5)
    base = gp.scenes.test_integrator(96, 64, spp=(3, 3), maxDepth=6)
    integ = P.NewPath(6, base.GetCamera(), base.GetSampler(), None, 1, P.Power)
    film, st, ofilm, ost = _render_both(gp, dev, scene, integ, 8)
    _assert_film_equal(film, ofilm, st, ost, "Path / Power")
    assert st["shadow_rays"] == 0 and st["closest_rays"] > st["camera_rays"] and not film[..., :3].any()
    uni = P.NewPath(6, base.GetCamera(), base.GetSampler(), None, 1, P.Uniform)
    film_u, st_u, _, _ = _render_both(gp, dev, scene, uni, 8)
    assert st_u["shadow_rays"] > 0 and film_u[..., :3].any()
    g = P.GpuScene(dev, scene)
    with pytest.raises(RuntimeError, match="rc=5"):
        P.Render(g, P.NewPath(6, base.GetCamera(), base.GetSampler(), None, 1, P.Spatial), 8)
    g.close()


def test_two_rank_partition_sums_to_single(gp, dev):
    # multi-GPU path emulated on one device: rank films summed == single-rank film up to summation order (1e-12 relative)
    scene, integ = gp.scenes.config1(W=160, H=90)
    g = gp.pbrt.GpuScene(dev, scene)
    gp.pbrt.Render(g, integ, 1)
    single = integ.GetCamera().GetFilm().pixels.copy()
    acc = np.zeros_like(single)
    rays = 0
    for r in range(2):
        st = gp.pbrt.Render(g, integ, 1, rank=r, world=2)
        acc += integ.GetCamera().GetFilm().pixels
        rays += st["camera_rays"]
    assert rays == 160 * 90 * 15
    assert np.array_equal(acc[..., 3], single[..., 3])
    assert np.allclose(acc, single, rtol=1e-12, atol=0)
    g.close()


def test_full_size_properties_config1(gp, dev):
    # BASELINE size (1920x1080, 15 effective spp): size-independent properties
    scene, integ = gp.scenes.config1()
    g = gp.pbrt.GpuScene(dev, scene)
    st = gp.pbrt.Render(g, integ, 1, flags=gp.abi.FLAG_COUNT_TRAVERSAL)
    film = integ.GetCamera().GetFilm().pixels
    W, H = 1920, 1080
    assert st["camera_rays"] == W * H * 15  # spp-1 samples per pixel
    # box filter r=1 with pFilm on integer corners: pixel (x,y) collects the samples of pixels x..x+1, y..y+1
    w = film[..., 3]
    assert w[0, 0] == 60 and w[H - 1, W - 1] == 15 and w[H - 1, 0] == 30 and w[0, W - 1] == 30 and np.all(w[:-1, :-1] == 60)
    assert st["closest_rays"] >= st["camera_rays"] and st["nodes_visited"] > st["closest_rays"]
    assert st["efloat_panics"] == 0 and st["stack_overflows"] == 0
    assert np.isfinite(film).all() and (film[..., :3] >= 0).all()
    # idempotence: a second render of the same frame is bit-identical
    first = film.copy()
    gp.pbrt.Render(g, integ, 1)
    assert np.array_equal(first, integ.GetCamera().GetFilm().pixels)
    g.close()


def test_full_size_composition_invariants_config2(gp, dev):
    # BASELINE size (1920x1080, 63 effective spp, FAST and — smaller — STRICT): properties of the COMPOSED Path.Li loop that follow
    # from the reference's source and hold exactly in floating point, checked on the CUDA path alone (no oracle in the loop):
    #  (1) linearity in the emitted radiance: every light x 2 (a power of two: exact) => the film's X, Y, Z exactly x 2, the
    #      weights and every ray count unchanged (no path decision reads a radiance: Russian roulette reads beta, the MIS
    #      weight reads pdfs, IsBlack tests survive the scaling — path.go:84-153, integrator.go:79-195);
    #  (2) maxDepth 1: `bounces++` comes first and `bounces >= maxDepth` ends the path before any light is gathered
    #      (path.go:41,66; the `bounces == 0` emission branch is dead): a black film, one closest-hit query per camera ray, no shadow ray;
    #  (3) every matte albedo 0: f is black, so EstimateDirect makes no visibility test (integrator.go:111-126) and SampleF ends
    #      the path (path.go:92-95): a black film and no shadow ray, while glass still bounces.
    P = gp.pbrt

    def render(scene, integ, **kw):
        g = P.GpuScene(dev, scene)
        st = P.Render(g, integ, 1, **kw)
        film = integ.GetCamera().GetFilm().pixels.copy()
        g.close()
        return film, st

    for dims, kw in ((dict(), dict(mode=gp.abi.MODE_FAST)), (dict(W=480, H=270), dict(mode=gp.abi.MODE_STRICT))):
        scene, integ = gp.scenes.config2(**dims)
        f1, s1 = render(scene, integ, **kw)
        scene2, integ2 = gp.scenes.config2(**dims)
        for l in scene2.lights:
            l.LEmit = [2.0 * v for v in l.LEmit]
        f2, s2 = render(scene2, integ2, **kw)
        assert np.array_equal(f2[..., 3], f1[..., 3])
        assert np.array_equal(f2[..., :3], 2.0 * f1[..., :3]), "the film is not linear in the emitted radiance"
        assert f1[..., :3].max() > 0
        for k in ("camera_rays", "closest_rays", "shadow_rays", "dead_mis_rays", "shaded_lanes"):
            assert s1[k] == s2[k], k
        # ... and so is DirectLighting.Li, its specular chain unwound included (directlighting.go:62-104), with both light strategies
        if dims:
            for strat in (P.UniformSampleOne, P.UniformSampleAll):
                films = []
                for scale in (1.0, 2.0):
                    sc_, it_ = gp.scenes.config2(**dims)
                    for l in sc_.lights:
                        l.LEmit = [scale * v for v in l.LEmit]
                    dl = P.NewDirectLighting(strat, 5, it_.GetCamera(), it_.GetSampler(), None)
                    films.append(render(sc_, dl, **kw)[0])
                assert np.array_equal(films[1][..., :3], 2.0 * films[0][..., :3]) and films[0][..., :3].max() > 0
        # (2)
        scene3, integ3 = gp.scenes.config2(**dims)
        cam, smp = integ3.GetCamera(), integ3.GetSampler()
        shallow = P.NewPath(1, cam, smp, None, 1, P.Uniform)
        f3, s3 = render(scene3, shallow, **kw)
        assert not f3[..., :3].any() and np.array_equal(f3[..., 3], f1[..., 3])
        assert s3["closest_rays"] == s3["camera_rays"] == s1["camera_rays"] and s3["shadow_rays"] == 0
        # (3)
        scene4, integ4 = gp.scenes.config2(**dims)
        for prim in scene4.aggregate.primitives:
            m = prim.material
            if isinstance(m, P.MatteMaterial):
                m.Kd.value = [0.0, 0.0, 0.0]
        f4, s4 = render(scene4, integ4, **kw)
        assert not f4[..., :3].any() and np.array_equal(f4[..., 3], f1[..., 3])
        assert s4["shadow_rays"] == 0 and s4["camera_rays"] == s1["camera_rays"] and s4["closest_rays"] > s4["camera_rays"]


def test_random_sampler_and_thin_lens_film_bit_exact(gp, dev):
    # sampler.RandomSampler (random.go): every dimension from the RNG, so pFilm is jittered and the box-filter footprint
    # varies per sample; lensRadius > 0 exercises ConcentricSampleDisk in GenerateRayDifferential (camera.go:201-212)
    P = gp.pbrt
    scene = gp.scenes.mixed_test_scene(60, seed=3)
    integ = gp.scenes.test_integrator(80, 48, maxDepth=5)
    cam = integ.GetCamera()
    cam.lensRadius, cam.focalDistance = 0.15, 20.0
    integ.sampler = P.NewRandomSampler(6)
    for tile in (1, 8):
        film, st, ofilm, ost = _render_both(gp, dev, scene, integ, tile)
        _assert_film_equal(film, ofilm, st, ost, f"random sampler + lens tile={tile}")
        assert st["camera_rays"] == 80 * 48 * 5


def test_partial_spheres_and_disks_bit_exact(gp, dev):
    # zMin/zMax/phiMax clipping (sphere.go:111-135, disk.go:86-93): root switching, the shadowed-phi quirk,
    # annulus and sector disks — rays and a small film
    P = gp.pbrt
    zero = P.NewConstantFloatTexture(0.0)
    uv = P.NewCheckerboard2D(P.NewUvMapping2D(6.0, 6.0, 0.0, 0.0), P.NewConstantSpectrumTexture(P.NewSpectrum(0.8)),
                             P.NewConstantSpectrumTexture(P.NewRGBSpectrum(0.2, 0.5, 0.9)))
    m = P.NewMatteMaterial(uv, zero)
    prims = []
    rng = gp.scenes.RNG(21)
    for i in range(40):
        c = (-6 + 12 * rng.UniformFloat(), 0.5 + 4 * rng.UniformFloat(), -6 + 12 * rng.UniformFloat())
        r = 0.4 + 0.8 * rng.UniformFloat()
        s = P.Sphere("p", P.Translate(c), bool(i % 2), r, zMin=-r * rng.UniformFloat(), zMax=r * (0.2 + 0.8 * rng.UniformFloat()),
                     phiMax=60.0 + 300.0 * rng.UniformFloat())
        prims.append(P.NewGeometricPrimitive(s, m))
    for i in range(10):
        xf = P.NewTransform(P.Translate((-5 + 10 * rng.UniformFloat(), 0.2 + 3 * rng.UniformFloat(), -5 + 10 * rng.UniformFloat())).Mul(
            P.RotateX(90.0 * rng.UniformFloat())).Matrix)
        prims.append(P.NewGeometricPrimitive(P.NewDisk(xf, 0.1, 1.5, 0.5 * rng.UniformFloat(), 90.0 + 270.0 * rng.UniformFloat()), m))
    prims.append(P.NewGeometricPrimitive(P.NewDisk(P.NewTransform(P.RotateX(90).Matrix), 0.0, 30.0, 0, 360), P.NewMatteMaterial(
        P.NewConstantSpectrumTexture(P.NewSpectrum(0.5)), zero)))
    scene = P.NewScene(P.NewBVH(prims, 4, P.SplitSAH), [P.NewPoint(P.Translate((3.0, 9.0, 4.0)), None, P.NewSpectrum(80)),
                                                          P.NewDistant(P.Translate((0.0, 0.0, 0.0)), P.NewSpectrum(0.4), (1.0, 1.0, -0.5))])
    g = P.GpuScene(dev, scene)
    s = OracleScene(scene, 1)
    rs = np.random.default_rng(9)
    n = 100000
    o = rs.uniform(-8, 8, size=(n, 3)); o[:, 1] = rs.uniform(0, 7, size=n)
    d = rs.normal(size=(n, 3))
    G, O = g.Intersect(o, d), s.intersect(o, d)
    _cmp_closest(G, O, "partial shapes")
    assert np.count_nonzero(O[0] >= 0) > n // 10
    assert np.array_equal(g.IntersectP(o, d), s.intersect_p(o, d))
    g.close(); s.close()
    integ = gp.scenes.test_integrator(96, 60, spp=(2, 2), pos=(10.0, 6.0, 10.0), look=(0.0, 1.5, 0.0), maxDepth=5)
    film, st, ofilm, ost = _render_both(gp, dev, scene, integ, 1)
    _assert_film_equal(film, ofilm, st, ost, "partial shapes film")


def test_fast_mode_sample_partition_sums_to_single(gp, dev):
    # FAST mode splits the work by SAMPLE index (north star): rank films sum to the single-rank film up to summation order
    scene = gp.scenes.mixed_test_scene(50, seed=9)
    integ = gp.scenes.test_integrator(96, 54, spp=(3, 3), maxDepth=5)
    g = gp.pbrt.GpuScene(dev, scene)
    gp.pbrt.Render(g, integ, 1, mode=gp.abi.MODE_FAST)
    single = integ.GetCamera().GetFilm().pixels.copy()
    acc = np.zeros_like(single)
    paths = 0
    for r in range(4):
        st = gp.pbrt.Render(g, integ, 1, mode=gp.abi.MODE_FAST, rank=r, world=4)
        acc += integ.GetCamera().GetFilm().pixels
        paths += st["camera_rays"]
    assert paths == 96 * 54 * 8
    assert np.array_equal(acc[..., 3], single[..., 3])
    assert np.allclose(acc, single, rtol=1e-12, atol=0)
    g.close()


def test_fast_mode_lane_groups_sum_to_single(gp, dev):
    # FAST mode, several lanes per pixel tile (each takes every n-th sample): same paths, same weights, film equal to the
    # one-lane-per-tile film up to summation order; also combined with the 2-rank sample split
    scene, integ = gp.scenes.config2(W=96, H=54, spp=(4, 4))
    g = gp.pbrt.GpuScene(dev, scene)
    st1 = gp.pbrt.Render(g, integ, 1, mode=gp.abi.MODE_FAST, groups=1)
    single = integ.GetCamera().GetFilm().pixels.copy()
    for groups in (2, 5, 15, 40):
        st = gp.pbrt.Render(g, integ, 1, mode=gp.abi.MODE_FAST, groups=groups)
        film = integ.GetCamera().GetFilm().pixels
        assert st["camera_rays"] == st1["camera_rays"] == 96 * 54 * 15
        assert st["closest_rays"] == st1["closest_rays"] and st["shadow_rays"] == st1["shadow_rays"]
        assert st["lanes"] == 96 * 54 * min(groups, 15)
        assert np.array_equal(film[..., 3], single[..., 3])
        assert np.allclose(film, single, rtol=1e-12, atol=0)
    acc = np.zeros_like(single)
    for r in range(2):
        gp.pbrt.Render(g, integ, 1, mode=gp.abi.MODE_FAST, rank=r, world=2, groups=3)
        acc += integ.GetCamera().GetFilm().pixels
    assert np.array_equal(acc[..., 3], single[..., 3])
    assert np.allclose(acc, single, rtol=1e-12, atol=0)
    # a pass that does not hold all lanes at once (max_lanes splits tiles' groups over passes)
    gp.pbrt.Render(g, integ, 1, mode=gp.abi.MODE_FAST, groups=4, max_lanes=7001)
    film = integ.GetCamera().GetFilm().pixels
    assert np.array_equal(film[..., 3], single[..., 3]) and np.allclose(film, single, rtol=1e-12, atol=0)
    g.close()


@pytest.mark.parametrize("maxDepth", [5, 4, 1])
def test_direct_lighting_film_bit_exact(gp, dev, maxDepth):
    # integrator.DirectLighting (directlighting.go:62-104), UniformSampleOne: the specular-transmission chain through the
    # glass sphere of config 2 is a recursion in the reference and a chain of frames on the GPU — same additions and
    # multiplications in the same order, so STRICT films are bit-identical
    P = gp.pbrt
    scene, integ = gp.scenes.config2(W=96, H=54, spp=(3, 3))
    dl = P.NewDirectLighting(P.UniformSampleOne, maxDepth, integ.GetCamera(), integ.GetSampler(), None)
    film, st, ofilm, ost = _render_both(gp, dev, scene, dl, 1)
    _assert_film_equal(film, ofilm, st, ost, f"direct lighting maxDepth {maxDepth}")
    if maxDepth > 1:
        assert st["closest_rays"] > st["camera_rays"]  # the glass sphere was entered
    else:
        assert st["closest_rays"] == st["camera_rays"]


def test_direct_lighting_mixed_scene_tiles_and_fast_mode(gp, dev):
    P = gp.pbrt
    scene = gp.scenes.mixed_test_scene(120, seed=13)
    base = gp.scenes.test_integrator(96, 64, spp=(3, 3), maxDepth=5)
    dl = P.NewDirectLighting(P.UniformSampleOne, 6, base.GetCamera(), base.GetSampler(), None)
    film, st, ofilm, ost = _render_both(gp, dev, scene, dl, 8)
    _assert_film_equal(film, ofilm, st, ost, "direct lighting mixed tile 8", exact=False)
    film, st, ofilm, ost = _render_both(gp, dev, scene, dl, 1, mode=gp.abi.MODE_FAST, groups=1)
    _assert_film_equal(film, ofilm, st, ost, "direct lighting mixed fast", exact=False)


def test_direct_lighting_sample_all_film_bit_exact(gp, dev):
    # UniformSampleAllLights (integrator.go:23-46): the cloned per-tile samplers carry no sample arrays (pixel.go:34-42),
    # so it is one sample per light for every light — on the GPU one shadow segment per light, summed in light order
    P = gp.pbrt
    scene = gp.scenes.mixed_test_scene(120, seed=13)
    base = gp.scenes.test_integrator(96, 64, spp=(3, 3), maxDepth=5)
    dl = P.NewDirectLighting(P.UniformSampleAll, 5, base.GetCamera(), base.GetSampler(), None)
    assert len(scene.lights) > 1
    film, st, ofilm, ost = _render_both(gp, dev, scene, dl, 8)
    _assert_film_equal(film, ofilm, st, ost, "direct lighting sample-all mixed tile 8", exact=False)
    assert st["shadow_rays"] > st["closest_rays"]  # several segments per hit
    film, st, ofilm, ost = _render_both(gp, dev, scene, dl, 1, mode=gp.abi.MODE_FAST, groups=1)
    _assert_film_equal(film, ofilm, st, ost, "direct lighting sample-all mixed fast", exact=False)
    scene, integ = gp.scenes.config2(W=96, H=54, spp=(3, 3))  # one light: a single segment per lane, same collection path
    dl = P.NewDirectLighting(P.UniformSampleAll, 5, integ.GetCamera(), integ.GetSampler(), None)
    film, st, ofilm, ost = _render_both(gp, dev, scene, dl, 1)
    _assert_film_equal(film, ofilm, st, ost, "direct lighting sample-all config2")


def test_one_scene_many_render_shapes_reuses_the_workspace(gp, dev):
    # the per-scene workspace is kept between gopbrt_render calls: integrators, strategies, modes and film sizes taken in
    # turn on ONE scene handle must each still match a fresh handle bit for bit
    P = gp.pbrt
    scene, integ = gp.scenes.config2(W=48, H=27, spp=(3, 3))
    cam, smp = integ.GetCamera(), integ.GetSampler()
    variants = [("path", integ, dict()), ("one", P.NewDirectLighting(P.UniformSampleOne, 5, cam, smp, None), dict()),
                ("all", P.NewDirectLighting(P.UniformSampleAll, 5, cam, smp, None), dict()),
                ("path fast", integ, dict(mode=gp.abi.MODE_FAST, groups=2)),
                ("all fast", P.NewDirectLighting(P.UniformSampleAll, 4, cam, smp, None), dict(mode=gp.abi.MODE_FAST)),
                ("path again", integ, dict()),
                ("path counted", integ, dict(flags=gp.abi.FLAG_COUNT_TRAVERSAL | gp.abi.FLAG_TIME_KERNELS))]
    shared = P.GpuScene(dev, scene)
    for name, ig, kw in variants:
        P.Render(shared, ig, 1, **kw)
        a = ig.GetCamera().GetFilm().pixels.copy()
        fresh = P.GpuScene(dev, scene)
        P.Render(fresh, ig, 1, **kw)
        b = ig.GetCamera().GetFilm().pixels.copy()
        fresh.close()
        assert np.array_equal(a, b), name
    shared.close()


def test_concurrent_renders_on_one_device_context(gp, dev):
    # each gRPC request of the reference renders on its own goroutine with its own scene: two host threads,
    # two scene handles, one device context — both films must equal their serial renders
    import threading
    P = gp.pbrt
    jobs = []
    for k, (W, H) in enumerate([(64, 40), (48, 27), (80, 45)]):
        scene, integ = gp.scenes.config2(W=W, H=H, spp=(3, 3))
        g = P.GpuScene(dev, scene)
        P.Render(g, integ, 1)
        jobs.append((g, integ, integ.GetCamera().GetFilm().pixels.copy()))
    out, errs = [None] * len(jobs), []

    def work(i):
        try:
            g, integ, _ = jobs[i]
            for _ in range(3):
                P.Render(g, integ, 1)
            out[i] = integ.GetCamera().GetFilm().pixels.copy()
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for i, (g, integ, ref) in enumerate(jobs):
        assert np.array_equal(out[i], ref)
        g.close()


def test_two_full_size_scenes_share_one_context(gp, dev):
    # two scene handles with 1080p FAST-mode workspaces (tens of GB of lane state each under the automatic lane groups) on ONE
    # context, rendered from two host threads: the per-handle lane budget must leave room for both, and each film must equal
    # the same handle's serial render bit for bit
    import threading
    P = gp.pbrt
    jobs = []
    for cfg, spp in (("config1", (4, 4)), ("config2", (4, 4))):
        scene, integ = getattr(gp.scenes, cfg)(spp=spp)
        g = P.GpuScene(dev, scene)
        st = P.Render(g, integ, 1, mode=gp.abi.MODE_FAST)
        assert st["lanes"] >= 1920 * 1080 * 4
        jobs.append((g, integ, integ.GetCamera().GetFilm().pixels.copy()))
    out, errs = [None, None], []

    def work(i):
        try:
            g, integ, _ = jobs[i]
            for _ in range(2):
                P.Render(g, integ, 1, mode=gp.abi.MODE_FAST)
            out[i] = integ.GetCamera().GetFilm().pixels.copy()
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for i, (g, integ, ref) in enumerate(jobs):
        assert np.array_equal(out[i], ref)
        g.close()


def test_degenerate_renders_match_oracle(gp, dev):
    # empty aggregate, no lights, a 1x1 film, one sample per pixel (pixel.go:48-52 increments first: zero samples run),
    # maxDepth 1, a film smaller than one tile
    P = gp.pbrt
    scene, integ = gp.scenes.config2(W=8, H=5, spp=(2, 2))
    cam, smp = integ.GetCamera(), integ.GetSampler()
    empty = P.NewScene(P.NewBVH([], 4, P.SplitSAH), [])
    for what, sc, ig, tile, kw in [
        ("empty scene", empty, integ, 1, {}),
        ("empty scene direct", empty, P.NewDirectLighting(P.UniformSampleAll, 5, cam, smp, None), 4, {}),
        ("no lights", P.NewScene(scene.aggregate, []), integ, 1, {}),
        ("no lights direct", P.NewScene(scene.aggregate, []), P.NewDirectLighting(P.UniformSampleOne, 5, cam, smp, None), 1, {}),
        ("tile larger than film", scene, integ, 64, {}),
        ("maxDepth 1", scene, P.NewPath(1, cam, smp, None, 1, P.Uniform), 1, {}),
        ("fast tiny", scene, integ, 1, dict(mode=gp.abi.MODE_FAST, groups=1)),
    ]:
        film, st, ofilm, ost = _render_both(gp, dev, sc, ig, tile, **kw)
        _assert_film_equal(film, ofilm, st, ost, what)
    s1, i1 = gp.scenes.config2(W=1, H=1, spp=(2, 2))
    film, st, ofilm, ost = _render_both(gp, dev, s1, i1, 1)
    _assert_film_equal(film, ofilm, st, ost, "1x1 film")
    s2, i2 = gp.scenes.config2(W=6, H=4, spp=(1, 1))
    film, st, ofilm, ost = _render_both(gp, dev, s2, i2, 1)
    _assert_film_equal(film, ofilm, st, ost, "one sample per pixel")
    assert st["camera_rays"] == 0 and not film.any()
