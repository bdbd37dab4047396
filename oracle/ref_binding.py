"""ctypes binding over oracle/_ref/libagpt_ref.so -- TEST INFRASTRUCTURE.

The library is the reference's own CPU path tracer (compiled from /root/reference by
oracle/Makefile) behind the harness in oracle/ref_harness.cpp.  Only tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes
import os
from ctypes import POINTER, byref, c_float, c_int, c_longlong, c_uint, c_ulonglong, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libagpt_ref.so")

HIT_DTYPE = np.dtype([("found", np.uint32), ("prim", np.int32), ("tri", np.int32), ("t", np.float32)])

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle ref` where /root/reference is mounted")
        L = ctypes.CDLL(LIB_PATH)
        L.agpt_ref_scene_create.restype = c_void_p
        L.agpt_ref_render.restype = c_longlong
        L.agpt_ref_primary_hits.restype = c_longlong
        L.agpt_ref_trace_rays.restype = c_longlong
        L.agpt_ref_stream_seed.restype = c_uint
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(POINTER(c_float))


class RefScene:
    def __init__(self, config, level=0):
        self._h = c_void_p(lib().agpt_ref_scene_create(c_int(config), c_int(level)))
        if not self._h:
            raise RuntimeError("unknown configuration")

    def close(self):
        if self._h:
            lib().agpt_ref_scene_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def count_rays(self):
        """Prepend the do-nothing ray-counting primitive (timing scenes only: shifts prim ids)."""
        lib().agpt_ref_scene_count_rays(self._h)

    def counts(self):
        a, b = c_int(), c_int()
        lib().agpt_ref_scene_counts(self._h, byref(a), byref(b))
        return dict(prims=a.value, lights=b.value)

    def render(self, W, H, s0, ns, max_depth, depth_arg=0, threads=0, crop=None, out=None):
        """Accumulator-layout float4 image (row H-1-y), summed over samples [s0, s0+ns)."""
        if out is None:
            out = np.zeros((H, W, 4), np.float32)
        x0, y0, x1, y1 = crop if crop else (0, 0, W, H)
        threads = threads or os.cpu_count() or 1
        n = lib().agpt_ref_render(self._h, c_int(W), c_int(H), c_int(x0), c_int(y0), c_int(x1), c_int(y1), c_int(s0), c_int(ns),
                                  c_int(max_depth), c_int(depth_arg), c_int(threads), _fp(out))
        return out, n

    def primary_hits(self, W, H, sample, threads=0, want_rays=False):
        hits = np.zeros(W * H, HIT_DTYPE)
        rays = np.zeros((W * H, 8), np.float32) if want_rays else None
        st = (c_ulonglong * 3)()
        threads = threads or os.cpu_count() or 1
        mism = lib().agpt_ref_primary_hits(self._h, c_int(W), c_int(H), c_int(sample), c_int(threads), hits.ctypes.data_as(c_void_p),
                                           _fp(rays) if want_rays else None, st)
        stats = dict(interior=st[0], boxes=st[1], tris=st[2], walk_mismatches=mism)
        return (hits, stats, rays) if want_rays else (hits, stats)

    def trace_rays(self, rays7, any_hit=False, threads=0):
        rays7 = np.ascontiguousarray(rays7, np.float32).reshape(-1, 7)
        hits = np.zeros(len(rays7), HIT_DTYPE)
        st = (c_ulonglong * 3)()
        threads = threads or os.cpu_count() or 1
        mism = lib().agpt_ref_trace_rays(self._h, c_longlong(len(rays7)), _fp(rays7), c_int(1 if any_hit else 0), c_int(threads),
                                         hits.ctypes.data_as(c_void_p), st)
        return hits, dict(interior=st[0], boxes=st[1], tris=st[2], walk_mismatches=mism)

    def li_pixels(self, W, H, xs, ys, ss, max_depth, depth_arg=0, want_draws=False):
        xs = np.ascontiguousarray(xs, np.int32); ys = np.ascontiguousarray(ys, np.int32); ss = np.ascontiguousarray(ss, np.int32)
        out = np.zeros((len(xs), 3), np.float32)
        draws = np.zeros(len(xs), np.int32)
        ip = lambda a: a.ctypes.data_as(POINTER(c_int))
        lib().agpt_ref_li_pixels(self._h, c_int(W), c_int(H), c_int(len(xs)), ip(xs), ip(ys), ip(ss), c_int(max_depth), c_int(depth_arg),
                                 _fp(out), ip(draws) if want_draws else None)
        return (out, draws) if want_draws else out

    def camera(self):
        out = np.zeros(19, np.float32)
        lib().agpt_ref_camera_export(self._h, _fp(out))
        return out

    def prim_info(self, prim):
        kind = c_int(); counts = (c_int * 5)(); hm = c_int(); il = c_int()
        rc = lib().agpt_ref_prim_info(self._h, c_int(prim), byref(kind), counts, byref(hm), byref(il))
        assert rc == 0
        return dict(kind=kind.value, nodes=counts[0], tris=counts[1], verts=counts[2], normals=counts[3], uvs=counts[4],
                    has_material=bool(hm.value), is_light=bool(il.value))

    def bvh(self, prim):
        info = self.prim_info(prim)
        nodes = np.zeros((info["nodes"], 8), np.uint32)
        order = np.zeros(info["tris"], np.int32)
        n = lib().agpt_ref_bvh_export(self._h, c_int(prim), nodes.ctypes.data_as(c_void_p), order.ctypes.data_as(POINTER(c_int)))
        assert n == info["nodes"]
        return nodes, order

    def mesh_verts(self, prim):
        info = self.prim_info(prim)
        v = np.zeros((info["tris"], 9), np.float32)
        lib().agpt_ref_mesh_export(self._h, c_int(prim), _fp(v))
        return v

    def material(self, prim):
        out = np.zeros(20, np.float32)
        lib().agpt_ref_material_export(self._h, c_int(prim), _fp(out))
        return out


def ray_counts(reset=True):
    out = (c_ulonglong * 2)()
    lib().agpt_ref_ray_counts(out, c_int(1 if reset else 0))
    return dict(closest=out[0], any=out[1])


def resolve(accum_rgba, samples):
    """Accumulator::CopyToSurface arithmetic (the reference's own lin2rgb / rgb2uint) on a float4 buffer."""
    a = np.ascontiguousarray(accum_rgba, np.float32).reshape(-1, 4)
    out = np.zeros(len(a), np.uint32)
    lib().agpt_ref_resolve(_fp(a), c_longlong(len(a)), c_int(samples), out.ctypes.data_as(POINTER(c_uint)))
    return out.reshape(np.shape(accum_rgba)[:-1])


def probe_bounds(boxes6, rays7):
    boxes6 = np.ascontiguousarray(boxes6, np.float32).reshape(-1, 6)
    rays7 = np.ascontiguousarray(rays7, np.float32).reshape(-1, 7)
    n = len(boxes6)
    hit = np.zeros(n, np.int32); t = np.zeros(n, np.float32)
    lib().agpt_ref_probe_bounds(c_int(n), _fp(boxes6), _fp(rays7), hit.ctypes.data_as(POINTER(c_int)), _fp(t))
    return hit, t


def probe_bsdf(mat6, in14, skip_specular):
    mat6 = np.ascontiguousarray(mat6, np.float32)
    in14 = np.ascontiguousarray(in14, np.float32).reshape(-1, 14)
    out = np.zeros((len(in14), 12), np.float32)
    lib().agpt_ref_probe_bsdf(c_int(len(in14)), _fp(mat6), _fp(in14), c_int(1 if skip_specular else 0), _fp(out))
    return out


def probe_sphere_sample(in9):
    in9 = np.ascontiguousarray(in9, np.float32).reshape(-1, 9)
    out = np.zeros((len(in9), 8), np.float32)
    lib().agpt_ref_probe_sphere_sample(c_int(len(in9)), _fp(in9), _fp(out))
    return out


def probe_stream(pixel_index, sample, k):
    out = np.zeros(k, np.float32)
    lib().agpt_ref_probe_stream(c_uint(pixel_index), c_uint(sample), c_int(k), _fp(out))
    return out


def probe_draw_order(state=12345):
    out = (c_float * 4)()
    lib().agpt_ref_probe_draw_order(c_uint(state), out)
    return list(out)


def sizes():
    out = (c_int * 7)()
    lib().agpt_ref_sizes(out)
    names = ["float3", "BVHNode", "Ray", "Primitive", "index_type", "BSDF", "SurfaceInteraction"]
    return dict(zip(names, list(out)))
