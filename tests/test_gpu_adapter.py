"""GPU: the drop-in proven against the reference's own classes.

tests/adapter/cuda_pathtracer.h is the header a maintainer of ag-pathtracer would add: `CudaPathTracer :
Integrator` over the reference's Scene / BVHTriMesh / material / light / Camera objects (integrator.h:28-31,
scene.h:3-30, bvhtrimesh.h:157-210 ...), reaching the GPU only through include/agpt.h.  oracle/Makefile compiles
it against the unmodified headers under /root/reference into oracle/_ref/libagpt_ref_adapter.so.  Here reference
Scene objects go through the adapter to the GPU and the result is compared with the reference's CPU
PathTracer on the very same objects -- the boundary checked against integrator.h / myapp.cpp:163-175 themselves,
not against the host mirror."""
import ctypes
import os
from ctypes import POINTER, c_float, c_int, c_uint, c_void_p

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libagpt_ref_adapter.so")


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        pytest.skip("oracle/_ref/libagpt_ref_adapter.so not built (needs /root/reference at build time)")
    L = ctypes.CDLL(LIB)
    L.agpt_ref_scene_create.restype = c_void_p
    L.agpt_ref_render.restype = ctypes.c_longlong
    L.agpt_ref_adapter_error.restype = ctypes.c_char_p
    return L


def fp(a):
    return a.ctypes.data_as(POINTER(c_float))


# (config, level, W, H, spp): analytic scene, BVH meshes + spheres + NEE/MIS, plain mesh + thin lens, environment map, corner cases
CASES = [(1, 0, 160, 90, 6), (3, 3, 160, 90, 4), (6, 2, 128, 72, 4), (7, 0, 96, 96, 4), (8, 3, 160, 90, 4), (5, 3, 128, 72, 3)]
DEFAULTS = {1: (5, 0), 3: (8, 0), 5: (16, 4), 6: (5, 0), 7: (5, 0), 8: (6, 4)}


@pytest.mark.parametrize("cfg,level,W,H,spp", CASES)
def test_reference_scene_through_the_adapter_equals_reference_cpu(lib, cfg, level, W, H, spp):
    md, da = DEFAULTS[cfg]
    h = c_void_p(lib.agpt_ref_scene_create(c_int(cfg), c_int(level)))
    assert h
    gpu = np.zeros((H, W, 4), np.float32)
    rgb = np.zeros((H, W), np.uint32)
    # two successive Render calls accumulate like successive Ticks
    for s0, ns, before in ((0, spp // 2, 0), (spp // 2, spp - spp // 2, spp // 2)):
        rc = lib.agpt_ref_adapter_render(h, c_int(W), c_int(H), c_int(s0), c_int(ns), c_int(before), c_int(md), c_int(da), fp(gpu), rgb.ctypes.data_as(POINTER(c_uint)))
        assert rc == 0, lib.agpt_ref_adapter_error().decode()
    cpu = np.zeros((H, W, 4), np.float32)
    n = lib.agpt_ref_render(h, c_int(W), c_int(H), c_int(0), c_int(0), c_int(W), c_int(H), c_int(0), c_int(spp), c_int(md), c_int(da), c_int(os.cpu_count() or 1), fp(cpu))
    assert n == W * H * spp
    same = (bits(gpu[..., :3]) == bits(cpu[..., :3])) | (np.isnan(gpu[..., :3]) & np.isnan(cpu[..., :3]))
    exact = same.all(axis=-1).mean()
    print(f"cfg{cfg}: adapter(GPU) vs reference CPU: bit-identical pixels {exact:.5f}")
    assert exact >= 0.999
    if np.isfinite(cpu).all():
        err = float(np.sqrt(np.mean((gpu[..., :3].astype(np.float64) - cpu[..., :3].astype(np.float64)) ** 2)) / np.mean(np.abs(cpu[..., :3])))
        assert err <= 1e-3
    # the displayed image: adapter CopyToSurface == the reference's lin2rgb / rgb2uint on the same film
    want = np.zeros((H, W), np.uint32)
    lib.agpt_ref_resolve(fp(gpu), ctypes.c_longlong(W * H), c_int(spp), want.ctypes.data_as(POINTER(c_uint)))
    assert np.array_equal(rgb, want)
    lib.agpt_ref_scene_destroy(h)


def test_li_entry_point_through_the_adapter(lib):
    """Integrator::Li(ray, scene, depth) -- the virtual the reference calls per pixel (myapp.cpp:168) and for the debug
    click (:196-198) -- on the adapter vs PathTracer::Li, same rays, same generator state."""
    cfg, level = 3, 3
    md, da = DEFAULTS[cfg]
    h = c_void_p(lib.agpt_ref_scene_create(c_int(cfg), c_int(level)))
    rng = np.random.default_rng(3)
    n = 64
    uv = rng.random((n, 2)).astype(np.float32)
    g = np.zeros((n, 3), np.float32); c = np.zeros((n, 3), np.float32)
    rc = lib.agpt_ref_adapter_li(h, c_int(n), fp(uv), c_uint(0x2545F491), c_int(md), c_int(da), fp(g), fp(c))
    assert rc == 0, lib.agpt_ref_adapter_error().decode()
    assert np.array_equal(bits(g), bits(c))
    assert np.abs(c).sum() > 0
    lib.agpt_ref_scene_destroy(h)
