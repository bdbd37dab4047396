"""ctypes binding over oracle/libagpt_oracle.so (the plain-C++ restatement) -- TEST INFRASTRUCTURE."""
import ctypes
import os
from ctypes import POINTER, c_float, c_int, c_longlong, c_uint, c_ulonglong, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libagpt_oracle.so")
HIT_DTYPE = np.dtype([("found", np.uint32), ("prim", np.int32), ("tri", np.int32), ("t", np.float32)])
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle port`")
        L = ctypes.CDLL(LIB_PATH)
        L.agpt_oracle_scene_create.restype = c_void_p
        L.agpt_oracle_render.restype = c_longlong
        L.agpt_oracle_primary_hits.restype = c_longlong
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(POINTER(c_float))


class PortScene:
    """The restatement fed with the flattened tables of a host-mirror scene (HostScene.tables())."""

    def __init__(self, host_scene):
        self._keep = host_scene                      # the tables borrow the host scene's memory
        self._h = c_void_p(lib().agpt_oracle_scene_create(host_scene.tables()))

    def close(self):
        if self._h:
            lib().agpt_oracle_scene_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render(self, W, H, s0, ns, max_depth, depth_arg=0, threads=0, crop=None, out=None, sample_stride=1):
        if out is None:
            out = np.zeros((H, W, 4), np.float32)
        x0, y0, x1, y1 = crop if crop else (0, 0, W, H)
        threads = threads or os.cpu_count() or 1
        cnt = (c_ulonglong * 6)()
        n = lib().agpt_oracle_render(self._h, c_int(W), c_int(H), c_int(x0), c_int(y0), c_int(x1), c_int(y1), c_int(s0), c_int(ns),
                                     c_int(sample_stride), c_int(max_depth), c_int(depth_arg), c_int(threads), _fp(out), cnt)
        names = ["rays_closest", "rays_any", "interior", "boxes", "tris", "analytic"]
        return out, dict(zip(names, list(cnt)), paths=n)

    def primary_hits(self, W, H, sample, threads=0):
        hits = np.zeros(W * H, HIT_DTYPE)
        lib().agpt_oracle_primary_hits(self._h, c_int(W), c_int(H), c_int(sample), c_int(threads or os.cpu_count() or 1), hits.ctypes.data_as(c_void_p))
        return hits

    def li_pixels(self, W, H, xs, ys, ss, max_depth, depth_arg=0):
        xs = np.ascontiguousarray(xs, np.int32); ys = np.ascontiguousarray(ys, np.int32); ss = np.ascontiguousarray(ss, np.int32)
        out = np.zeros((len(xs), 3), np.float32); draws = np.zeros(len(xs), np.int32)
        ip = lambda a: a.ctypes.data_as(POINTER(c_int))
        lib().agpt_oracle_li_pixels(self._h, c_int(W), c_int(H), c_int(len(xs)), ip(xs), ip(ys), ip(ss), c_int(max_depth), c_int(depth_arg), _fp(out), ip(draws))
        return out, draws


def probe_stream(pixel, sample, k):
    out = np.zeros(k, np.float32)
    lib().agpt_oracle_probe_stream(c_uint(pixel), c_uint(sample), c_int(k), _fp(out))
    return out
