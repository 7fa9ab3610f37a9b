"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).

Importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs; never from
the vecchio_b200 package.
"""
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from vecchio_b200 import _abi  # noqa: E402  (struct layouts of include/vecchio_gpu.h)


class orc_stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("paths", "rays", "dropped_samples", "n_node", "n_sph_rej", "n_sph_acc", "n_msph", "n_rect_rej",
                 "n_rect_acc", "n_box", "n_translate", "n_rotate", "n_medium", "n_texel", "n_perlin", "n_diffuse",
                 "n_dielectric", "n_metal", "n_emit_or_miss", "n_light_pdf", "rays_live")] + [("seconds", C.c_double), ("threads", C.c_int)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} missing: run `make oracle`")
        L = C.CDLL(path)
        vp = C.c_void_p
        L.orc_scene_create.argtypes = [C.POINTER(_abi.vk_scene_desc), C.POINTER(vp)]
        L.orc_scene_create.restype = C.c_int
        L.orc_scene_free.argtypes = [vp]
        L.orc_scene_free.restype = None
        L.orc_last_error.restype = C.c_char_p
        L.orc_intersect.argtypes = [vp, vp, C.c_size_t, vp, vp]
        L.orc_intersect.restype = C.c_int
        L.orc_render.argtypes = [vp, C.POINTER(_abi.vk_camera), C.POINTER(_abi.vk_render_params), vp, vp,
                                 C.POINTER(orc_stats), C.c_int]
        L.orc_render.restype = C.c_int
        L.orc_harvest_rays.argtypes = [vp, C.POINTER(_abi.vk_camera), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64,
                                       C.c_size_t, vp]
        L.orc_harvest_rays.restype = C.c_size_t
        L.orc_eval_batch.argtypes = [vp, vp, C.c_size_t]
        L.orc_eval_batch.restype = C.c_int
        L.orc_kat.argtypes = [vp, C.c_char_p, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.c_int]
        L.orc_kat.restype = C.c_int
        _lib = L
    return _lib


class OracleScene:
    """The reference's object graph rebuilt from a lowered scene (vecchio_b200.Scene)."""

    def __init__(self, scene):
        L = lib()
        h = C.c_void_p()
        rc = L.orc_scene_create(scene.desc_ptr, C.byref(h))
        if rc != 0:
            raise RuntimeError(L.orc_last_error().decode())
        self._h = h
        self._scene = scene  # keeps the texel/desc memory alive only during create; kept for convenience

    def intersect(self, rays, medium_xi=None):
        rays = np.ascontiguousarray(rays, dtype=_abi.RAY_DTYPE)
        out = np.zeros(len(rays), dtype=_abi.HIT_DTYPE)
        xi = None if medium_xi is None else np.ascontiguousarray(medium_xi, dtype=np.float32)
        rc = lib().orc_intersect(self._h, rays.ctypes.data, len(rays), xi.ctypes.data if xi is not None else None,
                                 out.ctypes.data)
        assert rc == 0
        return out

    def render(self, cam, params, want_sumsq=False, threads=0):
        n = params.width * params.height * 3
        rgb = np.empty(n, dtype=np.float32)
        sq = np.empty(n, dtype=np.float32) if want_sumsq else None
        st = orc_stats()
        rc = lib().orc_render(self._h, C.byref(cam), C.byref(params), rgb.ctypes.data,
                              sq.ctypes.data if want_sumsq else None, C.byref(st), threads)
        assert rc == 0
        shape = (params.height, params.width, 3)
        return rgb.reshape(shape), (sq.reshape(shape) if want_sumsq else None), st

    def harvest_rays(self, cam, width, height, max_depth, seed, max_rays):
        out = np.zeros(max_rays, dtype=_abi.RAY_DTYPE)
        n = lib().orc_harvest_rays(self._h, C.byref(cam), width, height, max_depth, seed, max_rays, out.ctypes.data)
        return out[:n]

    def eval_batch(self, recs):
        """The oracle's side of ``vk_eval_batch`` (same records, same variates)."""
        out = np.ascontiguousarray(recs, dtype=_abi.EVAL_DTYPE).copy()
        rc = lib().orc_eval_batch(self._h, out.ctypes.data, len(out))
        assert rc == 0, rc
        return out

    def kat(self, name, values, n_out):
        return kat(name, values, n_out, self._h)

    def close(self):
        if getattr(self, "_h", None):
            lib().orc_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def kat(name, values, n_out, scene_handle=None):
    vin = (C.c_float * len(values))(*[float(v) for v in values])
    vout = (C.c_float * max(n_out, 16))()
    n = lib().orc_kat(scene_handle, name.encode(), vin, len(values), vout, max(n_out, 16))
    if n < 0:
        raise ValueError(f"orc_kat({name}) rejected its arguments")
    return np.array(vout[:n_out], dtype=np.float32)
