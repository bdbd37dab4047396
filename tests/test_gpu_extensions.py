"""GPU: the extensions BASELINE's configurations name but the reference lacks.

PARITY UNPINNED BY THE REFERENCE: upstream has neither instancing (scene.h:5-28: a flat list, transforms baked into
vertices) nor a Russian roulette keyed on the bounce index (integrator.h:180 uses Li's constant depth argument).  What
each extension means is stated on the CPU in oracle/agpt_oracle.cpp first; these tests hold the CUDA path to that
statement bit for bit, and check the properties that tie the extension back to reference behaviour."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture(scope="module")
def port():
    from oracle import port_binding
    return port_binding


@pytest.mark.parametrize("level,W,H", [(3, 192, 108), (5, 160, 90)])
def test_instanced_scene_matches_cpu_statement(agpt, port, gpu_ctx, level, W, H):
    """cfg 9 = BASELINE config 4 with true instances: one mesh, eight placements (one rotated and scaled)."""
    d = agpt.config_defaults(9)
    hs = agpt.HostScene(9, level); ps = port.PortScene(hs)
    c = hs.counts()
    assert c["tris"] == 20 * 4 ** level, "one shared mesh, not eight copies"
    hs.upload(gpu_ctx); gpu_ctx.set_film(W, H); gpu_ctx.clear(); gpu_ctx.reset_stats()
    got = gpu_ctx.trace_primary(0, agpt.FLAG_COUNTERS)
    want = ps.primary_hits(W, H, 0)
    for f in ("found", "prim", "tri"):
        assert np.array_equal(got[f], want[f]), f
    assert np.array_equal(bits(got["t"]), bits(want["t"]))
    assert len(np.unique(got["prim"][got["found"] == 1])) >= 9, "every placement is seen"
    gpu_ctx.reset_stats()
    gpu_ctx.render(0, 4, d["max_depth"], d["depth_arg"], agpt.FLAG_COUNTERS)
    acc = gpu_ctx.read_accum()
    st = gpu_ctx.stats()
    ref, cnt = ps.render(W, H, 0, 4, d["max_depth"], d["depth_arg"])
    exact = np.all(bits(acc[..., :3]) == bits(ref[..., :3]), axis=-1).mean()
    print(f"cfg9 level {level}: bit-identical pixels {exact:.5f}")
    assert exact >= 0.999
    assert st.rays_shadow == cnt["rays_any"]
    assert st.rays_closest + st.rays_mis + st.rays_mis_culled + st.rays_tail_culled == cnt["rays_closest"]
    # the exact-filtered slab test takes the reference arithmetic's decisions in object space too
    gpu_ctx.clear(); gpu_ctx.render(0, 4, d["max_depth"], d["depth_arg"], agpt.FLAG_STRICT_BOXES)
    assert np.array_equal(bits(gpu_ctx.read_accum()), bits(acc))


def test_translated_instances_agree_with_baked_geometry(agpt, gpu_ctx):
    """Seven of cfg 9's placements are pure translations of cfg 4's spheres: same primitives hit, t within rounding."""
    level, W, H = 4, 192, 108
    hits = {}
    for cfg in (4, 9):
        hs = agpt.HostScene(cfg, level)
        hs.upload(gpu_ctx); gpu_ctx.set_film(W, H)
        hits[cfg] = gpu_ctx.trace_primary(0).copy()
    not5 = (hits[4]["prim"] != 6) & (hits[9]["prim"] != 6)              # primitive 6 = the rotated, scaled placement
    same = hits[4]["prim"][not5] == hits[9]["prim"][not5]
    assert same.mean() > 0.999
    dt = np.abs(hits[4]["t"][not5][same] - hits[9]["t"][not5][same])
    assert dt.max() < 1e-4


def test_hand_built_instance_with_shear_and_scale(agpt, port):
    """A non-uniformly scaled, rotated placement of a hand-built mesh through the raw C ABI: closest-hit, any-hit
    and visit counts against the CPU statement on the same tables."""
    from tests.test_gpu_corner_cases import chain_mesh
    mesh = chain_mesh(agpt, 33)
    inst = np.zeros(2, np.dtype([("mesh", np.int32), ("pad", np.int32, 3), ("o2w", np.float32, 12), ("w2o", np.float32, 12)]))
    def place(k, m3, t):
        m3 = np.asarray(m3, np.float64); t = np.asarray(t, np.float64)
        inv = np.linalg.inv(m3)
        inst[k]["mesh"] = 0
        inst[k]["o2w"] = np.concatenate([m3, t[:, None]], 1).astype(np.float32).ravel()
        inst[k]["w2o"] = np.concatenate([inv, (-inv @ t)[:, None]], 1).astype(np.float32).ravel()
    ca, sa = np.cos(.7), np.sin(.7)
    place(0, [[2 * ca, 0, .5 * sa], [.3, 1.5, 0], [-2 * sa, 0, .5 * ca]], [1, .5, -2])
    place(1, np.eye(3) * .25, [-3, 0, 1])
    prims = np.zeros(3, agpt.PRIM_DTYPE)
    prims[0] = (agpt.PRIM_INSTANCE, 0, 0, -1); prims[1] = (agpt.PRIM_BVH_MESH, 0, 0, -1); prims[2] = (agpt.PRIM_INSTANCE, 1, 0, -1)
    mat = agpt.make_material(agpt.MAT_DISNEY, (.7, .7, .7), .5, 0.)
    ctx = agpt.Context(0)
    ctx.upload_meshes([mesh])
    ctx.upload_table("spheres", np.zeros(0, agpt.SPHERE_DTYPE)); ctx.upload_table("planes", np.zeros(0, agpt.PLANE_DTYPE))
    ctx.upload_table("materials", [mat]); ctx.upload_table("lights", np.zeros(0, agpt.LIGHT_DTYPE))
    ctx.upload_instances(inst)
    ctx.upload_table("primitives", prims)
    ps = port.PortScene.from_tables(prims, [mesh], materials=[mat], instances=inst)
    rng = np.random.default_rng(9)
    n = 20000
    o = rng.uniform(-8, 8, (n, 3)); tgt = rng.uniform(-3, 5, (n, 3)) * np.array([1, .4, 1])
    rays = np.concatenate([o, tgt - o, np.where(rng.random(n) < .5, 3e38, rng.uniform(2, 12, n))[:, None]], 1).astype(np.float32)
    for any_hit in (False, True):
        ctx.reset_stats()
        got = ctx.trace_rays(rays, any_hit=any_hit, flags=agpt.FLAG_COUNTERS)
        want, cnt = ps.trace_rays(rays, any_hit=any_hit)
        assert np.array_equal(got["found"], want["found"])
        if not any_hit:
            assert np.array_equal(got["prim"], want["prim"]) and np.array_equal(got["tri"], want["tri"])
            assert np.array_equal(bits(got["t"]), bits(want["t"]))
            st = ctx.stats()
            assert (st.node_visits[0], st.box_tests[0], st.tri_tests[0]) == (cnt["interior"], cnt["boxes"], cnt["tris"])
            assert set(np.unique(got["prim"])) >= {0, 1, 2}
    ctx.close()


def test_russian_roulette_by_bounce(agpt, port, gpu_ctx):
    """AGPT_FLAG_RR_BY_BOUNCE: the roulette is live once bounces > 3 whatever Li's depth argument is."""
    cfg, level, W, H, spp = 5, 3, 160, 90, 6
    d = agpt.config_defaults(cfg)
    hs = agpt.HostScene(cfg, level); ps = port.PortScene(hs)
    hs.upload(gpu_ctx); gpu_ctx.set_film(W, H)

    def gpu(depth, depth_arg, flags):
        gpu_ctx.clear(); gpu_ctx.render(0, spp, depth, depth_arg, flags)
        return gpu_ctx.read_accum()

    by_bounce = gpu(16, 0, agpt.FLAG_RR_BY_BOUNCE)
    ps.set_rr_by_bounce(True)
    want, _ = ps.render(W, H, 0, spp, 16, 0)
    ps.set_rr_by_bounce(False)
    assert np.all(bits(by_bounce[..., :3]) == bits(want[..., :3]), axis=-1).mean() >= 0.999
    no_rr = gpu(16, 0, 0)
    ref_rr = gpu(16, 4, 0)                      # the reference's rule with its depth argument at 4: live from the first bounce
    assert not np.array_equal(bits(by_bounce), bits(no_rr)) and not np.array_equal(bits(by_bounce), bits(ref_rr))
    # paths that end before the rule can fire are untouched by it: up to max_depth 4 the films are the no-roulette films
    assert np.array_equal(bits(gpu(4, 0, agpt.FLAG_RR_BY_BOUNCE)), bits(gpu(4, 0, 0)))
    # single paths too
    xs = np.arange(0, W, 7, dtype=np.int32); ys = (xs * 3 % H).astype(np.int32); ss = (xs % 5).astype(np.int32)
    li = gpu_ctx.li_pixels(xs, ys, ss, 16, 0, agpt.FLAG_RR_BY_BOUNCE)
    ps.set_rr_by_bounce(True)
    want_li, _ = ps.li_pixels(W, H, xs, ys, ss, 16, 0)
    assert np.all(bits(li) == bits(want_li), axis=1).mean() >= 0.97


def test_glass_bsdf_against_reference_classes_and_cpu_statement(agpt, port, gpu_ctx):
    """Rough dielectric: the reflection lobe against golden vectors from the reference's own MicrofacetReflection +
    TrowbridgeReitzDistribution + FresnelDielectric; reflection + transmission against the CPU statement."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "glass_reflect.npz"))
    exact = total = 0
    for mat6, want in zip(g["mats"], g["out"]):
        m = agpt.make_material(agpt.MAT_GLASS, mat6[1:4], float(mat6[4]), float(mat6[5]))
        full = port.probe_bsdf(m, g["in"], False)
        got_full = gpu_ctx.probe_bsdf(m, g["in"], False)
        assert np.array_equal(bits(got_full[:, :4]), bits(full[:, :4])), "f / Pdf with the transmission lobe"
        assert np.isclose(got_full[:, 4:11], full[:, 4:11], rtol=2e-6, atol=1e-7).all()
        exact += int(np.all(bits(got_full) == bits(full), axis=1).sum()); total += len(full)
        m.lobes = agpt.LOBE_GLASS_REFLECT
        got = gpu_ctx.probe_bsdf(m, g["in"], False)
        assert np.array_equal(bits(got[:, :4]), bits(want[:, :4])), "reflection lobe: f / Pdf of the reference's classes"
        assert np.isclose(got[:, 4:11], want[:, 4:11], rtol=2e-6, atol=1e-7).all()
        exact += int(np.all(bits(got) == bits(want), axis=1).sum()); total += len(want)
    print(f"glass probe rows bit-identical: {exact}/{total}")
    assert exact / total > 0.995


def test_rough_glass_scene_matches_cpu_statement(agpt, port, gpu_ctx):
    """cfg 10 = BASELINE config 5 as worded: rough glass + diffuse interreflection, 16 bounces, roulette by bounce."""
    level, W, H, spp = 3, 192, 108, 6
    d = agpt.config_defaults(10)
    hs = agpt.HostScene(10, level); ps = port.PortScene(hs)
    ps.set_rr_by_bounce(True)
    hs.upload(gpu_ctx); gpu_ctx.set_film(W, H); gpu_ctx.clear(); gpu_ctx.reset_stats()
    gpu_ctx.render(0, spp, d["max_depth"], 0, agpt.FLAG_RR_BY_BOUNCE | agpt.FLAG_COUNTERS)      # (counting mode: exact ray accounting)
    got = gpu_ctx.read_accum()
    st = gpu_ctx.stats()
    want, cnt = ps.render(W, H, 0, spp, d["max_depth"], 0)
    exact = np.all(bits(got[..., :3]) == bits(want[..., :3]), axis=-1).mean()
    print(f"cfg10: bit-identical pixels {exact:.5f}")
    assert exact >= 0.999
    assert st.rays_shadow == cnt["rays_any"]
    assert st.rays_closest + st.rays_mis + st.rays_mis_culled + st.rays_tail_culled == cnt["rays_closest"]
    assert np.isfinite(got).all() and got[..., :3].min() >= 0
