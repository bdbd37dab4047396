"""Generate the golden fixtures under tests/golden/ from the REAL reference (oracle/_ref, i.e.
/root/reference compiled by oracle/Makefile).  Run here, where /root/reference is mounted:

    make -C oracle ref && python tests/golden/make_golden.py

The fixtures pin (a) the plain-C++ restatement oracle/agpt_oracle.cpp, (b) the host mirror's
scene construction, (c) the GPU kernels, on boxes where /root/reference does not exist.
Everything is produced by calling reference code; nothing here computes expected values itself.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_binding as ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
# (config, icosphere level, W, H, spp) -- small enough to commit, every code path covered
CASES = [(1, 0, 64, 36, 4), (2, 3, 64, 36, 2), (3, 2, 64, 36, 2), (4, 2, 64, 36, 2), (5, 2, 64, 36, 2), (6, 2, 64, 36, 4), (7, 0, 48, 48, 4),
         (8, 3, 96, 54, 4)]
DEFAULTS = {1: (5, 0), 2: (1, 0), 3: (8, 0), 4: (8, 0), 5: (16, 4), 6: (5, 0), 7: (5, 0), 8: (6, 4)}   # max_depth, depth_arg (config_scenes.h)

# cfg 8: rays no camera produces -- along +x and -x through all 30 chain triangles (29 pending far
# children; ten exact-t ties), and through the triangle whose |ng|^2 underflows, once with the ray
# extent beyond it (TriangleIntersect drops it, TriangleIntersectP keeps it) and once short of it.
CFG8_RAYS = [[-1, .1, 5.9, 1, 0, 0, 3.0e38], [-1, -.3, 6.2, 1, 0, 0, 3.0e38], [1e18, .1, 5.9, -1, 0, 0, 3.0e38], [-1, .1, 5.9, 1, 1e-3, 0, 3.0e38],
             [2.5e-13, 2.5e-13, -1, 0, 0, 1, 3.0e38], [2.5e-13, 2.5e-13, -1, 0, 0, 1, 1.05], [2.5e-13, 2.5e-13, -1, 0, 0, 1, .9999],
             [-2.8, .4, -1, 0, 0, 1, 3.0e38]]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def scene_fixture(cfg, level, W, H, spp):
    rs = ref.RefScene(cfg, level)
    md, da = DEFAULTS[cfg]
    d = {"case": np.array([cfg, level, W, H, spp, md, da], np.int32), "camera": rs.camera()}
    hits, st, rays = rs.primary_hits(W, H, 0, want_rays=True)
    assert st["walk_mismatches"] == 0
    d["hits"] = hits
    d["rays"] = rays
    d["walk_stats"] = np.array([st["interior"], st["boxes"], st["tris"]], np.uint64)
    acc, _ = rs.render(W, H, 0, spp, md, da)
    d["accum"] = acc
    # single paths: radiance + RNG draw counts (draw-order conformance)
    rng = np.random.RandomState(cfg)
    xs = rng.randint(0, W, 48).astype(np.int32); ys = rng.randint(0, H, 48).astype(np.int32); ss = rng.randint(0, 1000, 48).astype(np.int32)
    if cfg == 8:
        # add paths whose radiance is inf or NaN (the samples myapp.cpp:169-172 zeroes), found by scanning the film
        gx, gy = np.meshgrid(np.arange(W, dtype=np.int32), np.arange(H, dtype=np.int32))
        gx, gy = gx.ravel(), gy.ravel()
        for s in range(12):
            scan = rs.li_pixels(W, H, gx, gy, np.full_like(gx, s), md, da)
            bad = np.flatnonzero(np.isnan(scan).any(1))
            bad = np.concatenate([bad, np.flatnonzero(np.isinf(scan).any(1) & ~np.isnan(scan).any(1))[:4]])
            xs = np.concatenate([xs, gx[bad]]); ys = np.concatenate([ys, gy[bad]]); ss = np.concatenate([ss, np.full(len(bad), s, np.int32)])
    li, draws = rs.li_pixels(W, H, xs, ys, ss, md, da, want_draws=True)
    d["li_xs"], d["li_ys"], d["li_ss"], d["li"], d["li_draws"] = xs, ys, ss, li, draws
    # scene construction: per primitive kind / material constants / BVH digests
    n = rs.counts()["prims"]
    kinds, mats, digests = [], [], []
    for p in range(n):
        info = rs.prim_info(p)
        kinds.append([info["kind"], info["nodes"], info["tris"], int(info["has_material"]), int(info["is_light"])])
        mats.append(rs.material(p))
        if info["kind"] == 2:
            nodes, order = rs.bvh(p)
            digests.append([p, sha(nodes), sha(order), sha(rs.mesh_verts(p))])
        elif info["kind"] == 3:
            digests.append([p, "", "", sha(rs.mesh_verts(p))])
    d["prim_kinds"] = np.array(kinds, np.int32)
    d["materials"] = np.array(mats, np.float32)
    d["bvh_digests"] = np.array(digests, dtype="U64") if digests else np.zeros((0, 4), dtype="U64")
    # shadow / arbitrary rays through Scene::Intersect and IntersectP
    o = rng.uniform(-6, 6, (256, 3)).astype(np.float32); o[:, 1] = np.abs(o[:, 1])
    dd = rng.normal(size=(256, 3)).astype(np.float32)
    tm = np.where(rng.rand(256) < 0.5, np.float32(3.4028235e38), rng.uniform(0.5, 20, 256)).astype(np.float32)
    r7 = np.concatenate([o, dd, tm[:, None]], 1).astype(np.float32)
    if cfg == 8:
        r7 = np.concatenate([np.array(CFG8_RAYS, np.float32), r7])
    d["probe_rays"] = r7
    d["probe_closest"], _ = rs.trace_rays(r7, any_hit=False)
    d["probe_any"], _ = rs.trace_rays(r7, any_hit=True)
    np.savez_compressed(os.path.join(OUT, f"scene_cfg{cfg}.npz"), **d)
    print(f"cfg{cfg}: {n} prims, hits found {hits['found'].mean():.2f}, accum mean {acc[..., :3].mean():.4f}")


def function_fixture():
    rng = np.random.RandomState(7)
    d = {}
    # Bounds::Intersect incl. axis-parallel rays, rays starting on a slab plane (0/0 -> NaN), flat boxes
    n = 4096
    lo = rng.uniform(-2, 2, (n, 3)).astype(np.float32)
    hi = lo + rng.uniform(0, 2, (n, 3)).astype(np.float32)
    flat = rng.rand(n) < 0.1
    hi[flat, rng.randint(0, 3, flat.sum())] = lo[flat, rng.randint(0, 3, flat.sum())]
    hi = np.maximum(lo, hi)
    o = rng.uniform(-4, 4, (n, 3)).astype(np.float32)
    dr = rng.normal(size=(n, 3)).astype(np.float32)
    dr /= np.linalg.norm(dr, axis=1, keepdims=True).astype(np.float32)
    ax = rng.rand(n) < 0.15
    dr[ax, rng.randint(0, 3, ax.sum())] = 0.0
    onp = rng.rand(n) < 0.1
    k = rng.randint(0, 3, n)
    o[onp, k[onp]] = lo[onp, k[onp]]
    inside = rng.rand(n) < 0.2
    o[inside] = ((lo[inside] + hi[inside]) * 0.5).astype(np.float32)
    tm = np.where(rng.rand(n) < 0.5, np.float32(3.4028235e38), rng.uniform(0.1, 8, n)).astype(np.float32)
    boxes = np.concatenate([lo, hi], 1).astype(np.float32)
    rays = np.concatenate([o, dr, tm[:, None]], 1).astype(np.float32)
    d["bounds_boxes"], d["bounds_rays"] = boxes, rays
    d["bounds_hit"], d["bounds_t"] = ref.probe_bounds(boxes, rays)
    # BSDF f / Pdf / Sample_f for the material families of the configs
    mats = np.array([[1, .8, .3, .2, .5, 0.], [1, .944, .776, .373, .3, 1.], [1, .5, .5, .5, 1., 0.], [1, .912, .914, .92, .6, .5],
                     [1, .2, .45, .7, .05, 0.], [2, .9, .9, .9, 0., 0.]], np.float32)
    m = 512
    def unit(v):
        return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)
    dpdu = rng.normal(size=(m, 3)).astype(np.float32)
    dpdv = rng.normal(size=(m, 3)).astype(np.float32)
    nrm = unit(np.cross(dpdu, dpdv))
    wo = unit(rng.normal(size=(m, 3)) + 1.5 * nrm)
    wi = unit(rng.normal(size=(m, 3)) + 1.0 * nrm)
    flip = rng.rand(m) < 0.15
    wi[flip] = -wi[flip]
    u = rng.rand(m, 2).astype(np.float32)
    in14 = np.concatenate([dpdu, dpdv, wo, wi, u], 1).astype(np.float32)
    d["bsdf_mats"], d["bsdf_in"] = mats, in14
    d["bsdf_out_skip"] = np.stack([ref.probe_bsdf(mm, in14, True) for mm in mats])
    d["bsdf_out_all"] = np.stack([ref.probe_bsdf(mm, in14, False) for mm in mats])
    # Sphere::Sample / Pdf: far (Taylor branch), near, inside
    c = rng.uniform(-3, 3, (m, 3)).astype(np.float32)
    r = rng.uniform(.1, 2, (m, 1)).astype(np.float32)
    dist = np.concatenate([rng.uniform(1.01, 3, m // 2), rng.uniform(20, 200, m // 4), rng.uniform(0, .99, m - m // 2 - m // 4)]).astype(np.float32)[:, None]
    refp = (c + unit(rng.normal(size=(m, 3))) * r * dist).astype(np.float32)
    in9 = np.concatenate([c, r, refp, rng.rand(m, 2).astype(np.float32)], 1).astype(np.float32)
    d["sphere_in"], d["sphere_out"] = in9, ref.probe_sphere_sample(in9)
    # RNG streams + draw order
    pix = np.array([0, 1, 17, 640 * 360 - 1, 1920 * 1080 - 1, 3840 * 2160 - 1], np.uint32)
    smp = np.array([0, 1, 255, 1023, 65535, 4000000000], np.uint32)
    d["stream_pixels"], d["stream_samples"] = pix, smp
    d["stream_floats"] = np.stack([ref.probe_stream(int(p), int(s), 16) for p, s in zip(pix, smp)])
    d["draw_order"] = np.array(ref.probe_draw_order(12345), np.float32)
    d["sizes"] = np.array(list(ref.sizes().values()), np.int32)
    np.savez_compressed(os.path.join(OUT, "functions.npz"), **d)
    print("functions.npz: bounds hit rate", d["bounds_hit"].mean())


def glass_reflect_fixture():
    """The reflection lobe of the rough-dielectric extension from the reference's own MicrofacetReflection +
    TrowbridgeReitzDistribution + FresnelDielectric (oracle/ref_harness.cpp RefGlassReflection): f, Pdf, Sample_f."""
    rng = np.random.RandomState(23)
    mats = np.array([[3, 1., 1., 1., .3, 1.5], [3, .9, .8, .7, .05, 1.33], [3, 1., 1., 1., .8, 2.4]], np.float32)      # {3, Kr, roughness, eta}
    m = 512
    def unit(v):
        return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)
    dpdu = rng.normal(size=(m, 3)).astype(np.float32)
    dpdv = rng.normal(size=(m, 3)).astype(np.float32)
    nrm = unit(np.cross(dpdu, dpdv))
    wo = unit(rng.normal(size=(m, 3)) + 1.5 * nrm)
    back = rng.rand(m) < 0.3
    wo[back] = -wo[back]                       # from inside the medium too
    wi = unit(rng.normal(size=(m, 3)) + 1.0 * nrm)
    flip = rng.rand(m) < 0.3
    wi[flip] = -wi[flip]
    u = rng.rand(m, 2).astype(np.float32)
    in14 = np.concatenate([dpdu, dpdv, wo, wi, u], 1).astype(np.float32)
    d = {"mats": mats, "in": in14, "out": np.stack([ref.probe_bsdf(mm, in14, False) for mm in mats])}
    np.savez_compressed(os.path.join(OUT, "glass_reflect.npz"), **d)
    print("glass_reflect.npz: nonzero f rows", (np.abs(d["out"][..., :3]).sum(-1) > 0).mean())


if __name__ == "__main__":
    assert ref.available(), "build oracle/_ref first (make -C oracle ref)"
    if sys.argv[1:] == ["glass"]:
        glass_reflect_fixture()
        sys.exit(0)
    only = [int(a) for a in sys.argv[1:]]          # e.g. `make_golden.py 8` regenerates scene_cfg8.npz alone
    for case in CASES:
        if not only or case[0] in only:
            scene_fixture(*case)
    if not only:
        function_fixture()
        glass_reflect_fixture()
