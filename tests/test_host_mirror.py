"""CPU: the host mirror of the reference's scene API (ag-pathtracer_b200/host/) builds the same
cameras, materials, meshes and -- node for node -- the same SAH BVH as the reference."""
import hashlib
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("cfg", [1, 2, 3, 4, 5, 6, 7, 8])
def test_mirror_matches_golden(agpt, cfg):
    g = np.load(os.path.join(GOLDEN, f"scene_cfg{cfg}.npz"))
    level = int(g["case"][1])
    hs = agpt.HostScene(cfg, level)
    assert np.array_equal(hs.camera().view(np.uint32), g["camera"].view(np.uint32)), "camera vectors"
    kinds = g["prim_kinds"]
    assert hs.counts()["prims"] == len(kinds)
    for p, (kind, nodes, tris, has_mat, is_light) in enumerate(kinds):
        info = hs.prim_info(p)
        assert (info["kind"], int(info["has_material"]), int(info["is_light"])) == (kind, has_mat, is_light)
        assert np.array_equal(hs.material(p).view(np.uint32), g["materials"][p].view(np.uint32)), f"material constants of prim {p}"
        if kind >= 2:
            assert (info["nodes"], info["tris"]) == (nodes, tris)
    for p, dn, do, dv in g["bvh_digests"]:
        p = int(p)
        assert sha(hs.mesh_verts(p)) == dv, "mesh vertices"
        if dn:
            nodes, order = hs.bvh(p)
            assert sha(nodes) == dn, "flattened BVH nodes differ from the reference's"
            assert sha(order) == do, "leaf order differs from the reference's primitives[]"


@pytest.mark.parametrize("cfg,level", [(2, 6), (3, 4), (5, 4), (6, 3)])
def test_mirror_matches_reference_live(agpt, ref, cfg, level):
    hs = agpt.HostScene(cfg, level); rs = ref.RefScene(cfg, level)
    assert np.array_equal(hs.camera().view(np.uint32), rs.camera().view(np.uint32))
    for p in range(rs.counts()["prims"]):
        ri = rs.prim_info(p)
        assert np.array_equal(hs.material(p).view(np.uint32), rs.material(p).view(np.uint32))
        if ri["kind"] == 2:
            hn, ho = hs.bvh(p); rn, ro = rs.bvh(p)
            assert hn.shape == rn.shape and np.array_equal(hn, rn) and np.array_equal(ho, ro)


def test_struct_sizes_match_reference():
    g = np.load(os.path.join(GOLDEN, "functions.npz"))
    float3, bvhnode = int(g["sizes"][0]), int(g["sizes"][1])
    assert (float3, bvhnode) == (16, 32)      # float4 accumulator stride, 2 x 128-bit node loads


def test_bvh_layout_invariants(agpt):
    """Root at 0, slot 1 unused, sibling pairs on even indices (one 64-byte line), 2N-1 nodes."""
    hs = agpt.HostScene(2, 4)
    nodes, order = hs.bvh(0)
    first = nodes[:, 6].view(np.int32); count = nodes[:, 7].view(np.int32)
    interior = np.flatnonzero(count == 0)
    interior = interior[interior != 1]
    assert np.all(first[interior] % 2 == 0) and np.all(first[interior] >= 2)
    leaves = np.flatnonzero(count > 0)
    assert count[leaves].sum() == len(order) == 20 * 4 ** 4
    assert sorted(order.tolist()) == list(range(len(order)))
    assert len(nodes) == 2 * len(order) - 1 + 1


def _all_bvhs(hs):
    out = []
    for p in range(hs.counts()["prims"]):
        if hs.prim_info(p)["kind"] == 2:
            out.append(hs.bvh(p))
    return out


def test_parallel_build_equals_serial_build(agpt):
    """forking the top of the SAH build must not change one byte of the arrays
    (large enough that the builder really forks: > 4096 triangles per range)."""
    threads0, cache0 = agpt.get_build_options()
    try:
        agpt.set_build_options(1, "")
        serial = _all_bvhs(agpt.HostScene(2, 7))
        for t in (2, 5, 16):
            agpt.set_build_options(t, "")
            par = _all_bvhs(agpt.HostScene(2, 7))
            assert len(par) == len(serial) > 0
            for (n0, o0), (n1, o1) in zip(serial, par):
                assert n0.shape == n1.shape and np.array_equal(n0, n1) and np.array_equal(o0, o1), f"{t} threads"
    finally:
        agpt.set_build_options(threads0, cache0)


def test_bvh_cache_round_trip_and_damage(agpt, tmp_path):
    threads0, cache0 = agpt.get_build_options()
    try:
        agpt.set_build_options(0, "")
        fresh = _all_bvhs(agpt.HostScene(3, 4))
        agpt.set_build_options(0, str(tmp_path))
        assert agpt.get_build_options()[1] == str(tmp_path)
        built = _all_bvhs(agpt.HostScene(3, 4))              # builds and writes the cache files
        files = sorted(tmp_path.glob("*.agbvh"))
        assert len(files) >= 1 and not list(tmp_path.glob("*.tmp*"))
        cached = _all_bvhs(agpt.HostScene(3, 4))             # reads them back
        assert sorted(tmp_path.glob("*.agbvh")) == files, "the mesh key must not depend on padding bytes: same scene, same files"
        for (n0, o0), (n1, o1), (n2, o2) in zip(fresh, built, cached):
            assert np.array_equal(n0, n1) and np.array_equal(o0, o1)
            assert np.array_equal(n0, n2) and np.array_equal(o0, o2)
        # a truncated and a bit-flipped-header file are rejected and rebuilt, never trusted
        data = files[0].read_bytes()
        files[0].write_bytes(data[: len(data) // 2])
        if len(files) > 1:
            d1 = bytearray(files[1].read_bytes()); d1[9] ^= 0xFF; files[1].write_bytes(bytes(d1))
        again = _all_bvhs(agpt.HostScene(3, 4))
        for (n0, o0), (n1, o1) in zip(fresh, again):
            assert np.array_equal(n0, n1) and np.array_equal(o0, o1)
        assert files[0].stat().st_size == len(data)          # and the damaged file was replaced
    finally:
        agpt.set_build_options(threads0, cache0)
