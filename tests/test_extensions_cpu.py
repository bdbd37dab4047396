"""CPU: the extensions' CPU statements (oracle/agpt_oracle.cpp) -- what can be pinned to the reference is, the rest is
checked for the properties that make it a sane extension.  PARITY UNPINNED BY THE REFERENCE for: instancing, the
transmission lobe of the rough dielectric, the bounce-indexed roulette (upstream has none of them)."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture(scope="module")
def port():
    from oracle import port_binding
    return port_binding


def glass_record(agpt, mat6, transmit):
    m = agpt.make_material(agpt.MAT_GLASS, mat6[1:4], float(mat6[4]), float(mat6[5]))
    if not transmit:
        m.lobes = agpt.LOBE_GLASS_REFLECT
    return m


def test_glass_reflection_lobe_is_the_references_own_classes(agpt, port):
    """AGPT_LOBE_GLASS_REFLECT == MicrofacetReflection(Kr, TrowbridgeReitzDistribution, FresnelDielectric(1, eta)) built from
    the reference's classes (golden: tests/golden/glass_reflect.npz, made by oracle/_ref): f, Pdf and Sample_f bit for bit."""
    g = np.load(os.path.join(GOLDEN, "glass_reflect.npz"))
    for mat6, want in zip(g["mats"], g["out"]):
        got = port.probe_bsdf(glass_record(agpt, mat6, transmit=False), g["in"], False)
        assert np.array_equal(bits(got), bits(want))


def test_glass_furnace(agpt, port):
    """A rough-glass sphere (Kr = Kt = 1) under a uniform white sky: reflection + transmission return what arrives
    (single scattering loses a little at high roughness, never gains)."""
    cam = agpt.HostScene(1, 0).camera()
    for rough, lo in ((0.05, 0.97), (0.3, 0.95)):
        mat = agpt.make_material(agpt.MAT_GLASS, (1, 1, 1), rough, 1.5)
        prims = np.zeros(1, agpt.PRIM_DTYPE); spheres = np.zeros(1, agpt.SPHERE_DTYPE); lights = np.zeros(1, agpt.LIGHT_DTYPE)
        spheres[0] = ((0, 0, 0), 1.0, 1.0, (0, 0, 0)); prims[0] = (agpt.PRIM_SPHERE, 0, 0, -1)
        lights[0] = (agpt.LIGHT_UNIFORM_INFINITE, -1, (0, 0), (1, 1, 1), 0)
        ps = port.PortScene.from_tables(prims, spheres=spheres, materials=[mat], lights=lights, camera=cam)
        W, H, spp = 64, 36, 48
        acc, _ = ps.render(W, H, 0, spp, 24, 0)
        centre = float(np.median(acc[H // 2 - 4:H // 2 + 4, W // 2 - 4:W // 2 + 4, :3]) / spp)
        assert lo < centre < 1.03, (rough, centre)


def test_transmission_pdf_matches_its_sampler(agpt, port):
    """Sample_f of the glass BSDF returns the pdf that Pdf() reports for the direction it sampled, and f() its value."""
    g = np.load(os.path.join(GOLDEN, "glass_reflect.npz"))
    mat = glass_record(agpt, g["mats"][0], transmit=True)
    a = g["in"].copy()
    s = port.probe_bsdf(mat, a, False)
    ok = s[:, 10] > 0
    assert ok.mean() > 0.5
    b = a[ok].copy()
    b[:, 9:12] = s[ok, 4:7]                        # evaluate at the sampled direction
    e = port.probe_bsdf(mat, b, False)
    assert np.allclose(e[:, 3], s[ok, 10], rtol=2e-4, atol=1e-6)
    assert np.allclose(e[:, :3], s[ok, 7:10], rtol=2e-4, atol=1e-6)
    # some samples cross the interface
    dpdu, dpdv = b[:, 0:3], b[:, 3:6]
    n = np.cross(dpdu, dpdv)
    crossed = np.sign((b[:, 6:9] * n).sum(1)) != np.sign((b[:, 9:12] * n).sum(1))
    assert 0.2 < crossed.mean() < 0.9


def test_instanced_scene_flattens_to_one_mesh(agpt, port):
    hs = agpt.HostScene(9, 3)
    c = hs.counts()
    assert c["tris"] == 1280 and c["prims"] == 11
    assert [hs.prim_info(p)["kind"] for p in range(1, 9)] == [agpt.PRIM_INSTANCE] * 8
    ps = port.PortScene(hs)
    hits = ps.primary_hits(96, 54, 0)
    assert len(np.unique(hits["prim"][hits["found"] == 1])) >= 8
    # identity placement == the plain mesh: same hits, same t bits
    from tests.test_gpu_corner_cases import chain_mesh
    mesh = chain_mesh(agpt, 17)
    inst = np.zeros(1, np.dtype([("mesh", np.int32), ("pad", np.int32, 3), ("o2w", np.float32, 12), ("w2o", np.float32, 12)]))
    eye = np.concatenate([np.eye(3), np.zeros((3, 1))], 1).astype(np.float32).ravel()
    inst[0]["o2w"] = eye; inst[0]["w2o"] = eye
    mat = agpt.make_material(agpt.MAT_DISNEY, (.7, .7, .7), .5, 0.)
    pa = np.zeros(1, agpt.PRIM_DTYPE); pa[0] = (agpt.PRIM_BVH_MESH, 0, 0, -1)
    pb = np.zeros(1, agpt.PRIM_DTYPE); pb[0] = (agpt.PRIM_INSTANCE, 0, 0, -1)
    rng = np.random.default_rng(1)
    o = rng.uniform(-4, 6, (2000, 3)); t = rng.uniform(-1, 3, (2000, 3)) * np.array([1, .5, .5])
    d = t - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)          # unit directions: the object-space ray of an identity placement is the world ray
    rays = np.concatenate([o, d, np.full((2000, 1), 3e38)], 1).astype(np.float32)
    a, _ = port.PortScene.from_tables(pa, [mesh], materials=[mat]).trace_rays(rays)
    b, _ = port.PortScene.from_tables(pb, [mesh], materials=[mat], instances=inst).trace_rays(rays)
    assert np.array_equal(a["found"], b["found"]) and np.array_equal(a["tri"], b["tri"]) and np.array_equal(bits(a["t"]), bits(b["t"]))


def test_roulette_by_bounce_leaves_short_paths_alone(agpt, port):
    hs = agpt.HostScene(5, 2)
    ps = port.PortScene(hs)
    base, _ = ps.render(64, 36, 0, 2, 4, 0)
    ps.set_rr_by_bounce(True)
    same, _ = ps.render(64, 36, 0, 2, 4, 0)
    deep, _ = ps.render(64, 36, 0, 2, 16, 0)
    ps.set_rr_by_bounce(False)
    deep_no, _ = ps.render(64, 36, 0, 2, 16, 0)
    assert np.array_equal(bits(base), bits(same))
    assert not np.array_equal(bits(deep), bits(deep_no))
