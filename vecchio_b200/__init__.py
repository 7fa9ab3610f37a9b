"""vecchio_b200 -- B200-native (sm_100a) implementation of vecchio's path-tracing sample loop.

Python is only the harness language here (tests, bench).  The product is two shared libraries:

* ``lib/libvecchio_gpu.so``  -- hand-written CUDA kernels behind the C ABI of ``include/vecchio_gpu.h``
* ``lib/libvecchio_host.so`` -- the reference's scene/Hittable/Material API on the host with ``lower()``

There is no CPU fallback: :class:`Context` raises if the CUDA library is missing or no device exists.
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import *  # noqa: F401,F403  (struct mirrors and constants)

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)
ASSETS_DIR = os.path.join(REPO_ROOT, "assets")
_LIBDIR = os.path.join(_HERE, "lib")

_host = None
_gpu = None


class VecchioError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def host_lib():
    """libvecchio_host.so (scene front end); built by `make host`."""
    global _host
    if _host is None:
        path = os.path.join(_LIBDIR, "libvecchio_host.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: run `make host` (or __graft_entry__.build())")
        L = C.CDLL(path)
        L.vkh_scene_build.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint32, C.POINTER(C.c_void_p)]
        L.vkh_scene_build.restype = C.c_int
        L.vkh_scene_free.argtypes = [C.c_void_p]
        L.vkh_scene_free.restype = None
        L.vkh_scene_desc.argtypes = [C.c_void_p]
        L.vkh_scene_desc.restype = C.POINTER(_abi.vk_scene_desc)
        L.vkh_scene_aspect_ratio.argtypes = [C.c_void_p]
        L.vkh_scene_aspect_ratio.restype = C.c_float
        L.vkh_scene_next_camera.argtypes = [C.c_void_p, C.POINTER(_abi.vk_camera)]
        L.vkh_scene_next_camera.restype = C.c_int
        L.vkh_camera_new.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float,
                                     C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(_abi.vk_camera)]
        L.vkh_camera_new.restype = None
        L.vkh_decode_png.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.vkh_decode_png.restype = C.c_long
        L.vkh_frame_to_rgb8.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.vkh_frame_to_rgb8.restype = None
        L.vkh_write_ppm.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.vkh_write_ppm.restype = C.c_int
        L.vkh_frame_filename.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_size_t]
        L.vkh_frame_filename.restype = C.c_int
        L.vkh_last_error.restype = C.c_char_p
        _host = L
    return _host


def gpu_lib():
    """libvecchio_gpu.so (CUDA kernels + C ABI); built by `make gpu`.  Raises if absent."""
    global _gpu
    if _gpu is None:
        path = os.environ.get("VECCHIO_GPU_LIB") or os.path.join(_LIBDIR, "libvecchio_gpu.so")  # env: tuning sweeps
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: run `make gpu` (or __graft_entry__.build()). "
                              "There is no CPU fallback for the render path.")
        L = C.CDLL(path)
        vp = C.c_void_p
        L.vk_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.vk_destroy.argtypes = [vp]
        L.vk_destroy.restype = None
        L.vk_last_error.argtypes = [vp]
        L.vk_last_error.restype = C.c_char_p
        L.vk_scene_upload.argtypes = [vp, C.POINTER(_abi.vk_scene_desc)]
        L.vk_render.argtypes = [vp, C.POINTER(_abi.vk_camera), C.POINTER(_abi.vk_render_params), vp, vp,
                                C.POINTER(_abi.vk_stats)]
        L.vk_scene_check.argtypes = [C.POINTER(_abi.vk_scene_desc), C.POINTER(_abi.vk_scene_info), C.c_char_p, C.c_size_t]
        L.vk_scene_check.restype = C.c_int
        L.vk_render_rgb8.argtypes = [vp, C.POINTER(_abi.vk_camera), C.POINTER(_abi.vk_render_params), vp, C.POINTER(_abi.vk_stats)]
        L.vk_render_device.argtypes = [vp, C.POINTER(_abi.vk_camera), C.POINTER(_abi.vk_render_params), vp, vp,
                                       C.POINTER(_abi.vk_stats)]
        L.vk_finalize_device.argtypes = [vp, vp, vp, C.c_size_t, C.c_uint32]
        L.vk_set_stream.argtypes = [vp, vp]
        L.vk_flush_stats.argtypes = [vp, C.POINTER(_abi.vk_stats)]
        L.vk_intersect.argtypes = [vp, vp, C.c_size_t, vp, C.c_uint32, vp]
        L.vk_eval_batch.argtypes = [vp, vp, C.c_size_t, C.c_uint32]
        L.vk_measure_peaks.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.vk_device_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, C.c_size_t]
        L.vk_multi_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
        L.vk_multi_destroy.argtypes = [vp]
        L.vk_multi_destroy.restype = None
        L.vk_multi_last_error.argtypes = [vp]
        L.vk_multi_last_error.restype = C.c_char_p
        L.vk_multi_device_count.argtypes = [vp]
        L.vk_multi_scene_upload.argtypes = [vp, C.POINTER(_abi.vk_scene_desc)]
        L.vk_multi_render.argtypes = [vp, C.POINTER(_abi.vk_camera), C.POINTER(_abi.vk_render_params), vp, vp, C.POINTER(_abi.vk_stats)]
        L.vk_multi_render_rgb8.argtypes = [vp, C.POINTER(_abi.vk_camera), C.POINTER(_abi.vk_render_params), vp, C.POINTER(_abi.vk_stats)]
        for f in ("vk_multi_create", "vk_multi_device_count", "vk_multi_scene_upload", "vk_multi_render", "vk_multi_render_rgb8"):
            getattr(L, f).restype = C.c_int
        for f in ("vk_create", "vk_scene_upload", "vk_render", "vk_render_rgb8", "vk_render_device", "vk_finalize_device",
                  "vk_set_stream", "vk_flush_stats", "vk_intersect", "vk_eval_batch", "vk_measure_peaks", "vk_device_info"):
            getattr(L, f).restype = C.c_int
        _gpu = L
    return _gpu


GPU_SYMBOLS = ["vk_create", "vk_destroy", "vk_last_error", "vk_scene_check", "vk_scene_upload", "vk_render", "vk_render_rgb8", "vk_render_device",
               "vk_finalize_device", "vk_set_stream", "vk_flush_stats", "vk_intersect", "vk_eval_batch", "vk_measure_peaks",
               "vk_device_info", "vk_multi_create", "vk_multi_destroy", "vk_multi_last_error", "vk_multi_device_count",
               "vk_multi_scene_upload", "vk_multi_render", "vk_multi_render_rgb8"]
HOST_SYMBOLS = ["vkh_scene_build", "vkh_scene_free", "vkh_scene_desc", "vkh_scene_aspect_ratio",
                "vkh_scene_next_camera", "vkh_camera_new", "vkh_decode_png", "vkh_frame_to_rgb8", "vkh_write_ppm",
                "vkh_frame_filename", "vkh_last_error"]


class Scene:
    """A scene built with the reference's scene API (src/scene.rs) and lowered for the GPU.

    ``Scene("cornell_box")`` mirrors ``cornell_box()`` + ``BVHNode::new(&mut config.world)``
    (src/main.rs:159-169).  ``seed`` stands in for the reference's unseeded thread_rng().
    """

    def __init__(self, name, seed=1, param=0, assets_dir=None):
        L = host_lib()
        h = C.c_void_p()
        rc = L.vkh_scene_build(name.encode(), seed, (assets_dir or ASSETS_DIR).encode(), param, C.byref(h))
        if rc != 0:
            raise VecchioError(rc, L.vkh_last_error().decode())
        self._h = h
        self.name = name
        self.desc_ptr = L.vkh_scene_desc(h)
        self.desc = self.desc_ptr.contents
        self.aspect_ratio = float(L.vkh_scene_aspect_ratio(h))

    def next_camera(self):
        """``config.cam_iter.next()`` (src/main.rs:176); None when exhausted."""
        cam = _abi.vk_camera()
        return cam if host_lib().vkh_scene_next_camera(self._h, C.byref(cam)) else None

    def height_for(self, width):
        """``((width as f32) / config.aspect_ratio) as usize`` (src/main.rs:172)."""
        return int(np.float32(width) / np.float32(self.aspect_ratio))

    def census(self):
        d = self.desc
        return {k: int(getattr(d, "n_" + k)) for k in
                ("nodes", "spheres", "mspheres", "rects", "boxes", "xforms", "media", "lights", "materials",
                 "textures", "perlins")} | {"texel_bytes": int(d.n_texel_bytes)}

    def nbytes(self):
        d = self.desc
        return (d.n_nodes * 32 + d.n_spheres * 20 + d.n_mspheres * 48 + d.n_rects * 32 + d.n_boxes * 32 +
                d.n_xforms * 32 + d.n_media * 16 + d.n_lights * 4 + d.n_materials * 16 + d.n_textures * 16 +
                int(d.n_texel_bytes) + d.n_perlins * C.sizeof(_abi.vk_perlin))

    def close(self):
        if getattr(self, "_h", None):
            host_lib().vkh_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def scene_check(desc_ptr):
    """``vk_scene_check``: validate + plan the device layout on the host (no GPU needed).  Returns a dict of
    what ``vk_scene_upload`` would build; raises VecchioError exactly where the upload would fail."""
    info = _abi.vk_scene_info()
    err = C.create_string_buffer(256)
    rc = gpu_lib().vk_scene_check(desc_ptr, C.byref(info), err, 256)
    if rc != 0:
        raise VecchioError(rc, err.value.decode())
    return {n: int(getattr(info, n)) for n, _ in info._fields_}


def camera_new(lookfrom, lookat, vup, vfov, aspect_ratio, aperture, focus_dist, time0, time1):
    """``Camera::new`` (src/main.rs:71-109)."""
    f3 = lambda v: (C.c_float * 3)(*[float(x) for x in v])  # noqa: E731
    cam = _abi.vk_camera()
    host_lib().vkh_camera_new(f3(lookfrom), f3(lookat), f3(vup), vfov, aspect_ratio, aperture, focus_dist, time0,
                              time1, C.byref(cam))
    return cam


def decode_png(path):
    """``ImageTexture::new`` (src/material.rs:269-279): (h, w, 3) uint8."""
    L = host_lib()
    w, h = C.c_uint32(), C.c_uint32()
    n = L.vkh_decode_png(path.encode(), None, 0, C.byref(w), C.byref(h))
    if n < 0:
        raise VecchioError(-1, L.vkh_last_error().decode())
    buf = np.empty(n, dtype=np.uint8)
    L.vkh_decode_png(path.encode(), buf.ctypes.data, n, C.byref(w), C.byref(h))
    return buf.reshape(h.value, w.value, 3)


def frame_to_rgb8(frame):
    """``vkh_frame_to_rgb8``: an (H, W, 3) linear frame with row 0 = bottom (what ``Context.render`` returns) ->
    the (H, W, 3) uint8 numbers of the reference's P3 file, top row first (src/main.rs:209-212)."""
    frame = np.ascontiguousarray(frame, dtype=np.float32)
    h, w, _ = frame.shape
    out = np.empty((h, w, 3), dtype=np.uint8)
    host_lib().vkh_frame_to_rgb8(frame.ctypes.data, w, h, out.ctypes.data)
    return out


def frame_filename(file_idx, out_dir=""):
    """``format!("output_{:04}.ppm", file_idx)`` (src/main.rs:201), under out_dir."""
    buf = C.create_string_buffer(4096)
    if host_lib().vkh_frame_filename(out_dir.encode(), file_idx, buf, 4096) < 0:
        raise ValueError("frame file name too long")
    return buf.value.decode()


def write_ppm(path, rgb8):
    """``vkh_write_ppm``: the P3 writer of src/main.rs:201-213; rgb8 = (H, W, 3) uint8 in file order (what
    ``Context.render_rgb8`` and ``frame_to_rgb8`` return)."""
    rgb8 = np.ascontiguousarray(rgb8, dtype=np.uint8)
    h, w, _ = rgb8.shape
    L = host_lib()
    if L.vkh_write_ppm(path.encode(), rgb8.ctypes.data, w, h) != 0:
        raise VecchioError(-1, L.vkh_last_error().decode())


def render_params(width, height, spp, max_depth=100, seed=1, spp_begin=0, spp_count=0, variant=0, flags=0,
                  background=(0.0, 0.0, 0.0)):
    p = _abi.vk_render_params()
    p.width, p.height, p.spp, p.spp_begin, p.spp_count, p.max_depth = width, height, spp, spp_begin, spp_count, max_depth
    p.seed = seed
    p.background[:] = background
    p.variant, p.flags = variant, flags
    return p


class Context:
    """One GPU context (``vk_create``).  Fails loudly without the CUDA library or a device."""

    def __init__(self, device=0):
        L = gpu_lib()
        h = C.c_void_p()
        rc = L.vk_create(device, C.byref(h))
        if rc != 0:
            raise VecchioError(rc, (L.vk_last_error(None) or b"").decode())
        self._h = h
        self._L = L
        self.device = device

    def _check(self, rc):
        if rc != 0:
            raise VecchioError(rc, (self._L.vk_last_error(self._h) or b"").decode())

    def upload(self, scene):
        self._check(self._L.vk_scene_upload(self._h, scene.desc_ptr))

    def render(self, cam, params, want_sumsq=False):
        """``vk_render``: the whole sample loop, host buffers in and out.  Returns (rgb, sumsq, stats)."""
        n = params.width * params.height * 3
        rgb = np.empty(n, dtype=np.float32)
        sq = np.empty(n, dtype=np.float32) if want_sumsq else None
        st = _abi.vk_stats()
        self._check(self._L.vk_render(self._h, C.byref(cam), C.byref(params), rgb.ctypes.data,
                                      sq.ctypes.data if want_sumsq else None, C.byref(st)))
        shape = (params.height, params.width, 3)
        return rgb.reshape(shape), (sq.reshape(shape) if want_sumsq else None), st

    def render_rgb8(self, cam, params):
        """``vk_render_rgb8``: the frame as the reference's PPM holds it -- (H, W, 3) uint8, row 0 = top,
        every channel through ``Vec3::to_color`` (src/main.rs:201-214).  Returns (rgb8, stats)."""
        out = np.empty(params.width * params.height * 3, dtype=np.uint8)
        st = _abi.vk_stats()
        self._check(self._L.vk_render_rgb8(self._h, C.byref(cam), C.byref(params), out.ctypes.data, C.byref(st)))
        return out.reshape(params.height, params.width, 3), st

    def render_device(self, cam, params, d_sum_ptr, d_sumsq_ptr=None, want_stats=True):
        """``vk_render_device``: per-pixel SUMS of an spp slice into device memory.  With
        ``want_stats=False`` the call only enqueues work (collect counts with flush_stats())."""
        st = _abi.vk_stats() if want_stats else None
        self._check(self._L.vk_render_device(self._h, C.byref(cam), C.byref(params), d_sum_ptr, d_sumsq_ptr,
                                             C.byref(st) if want_stats else None))
        return st

    def set_stream(self, stream_ptr):
        self._check(self._L.vk_set_stream(self._h, stream_ptr))

    def flush_stats(self):
        st = _abi.vk_stats()
        self._check(self._L.vk_flush_stats(self._h, C.byref(st)))
        return st

    def finalize_device(self, d_sum_ptr, d_rgb_ptr, n_floats, spp):
        self._check(self._L.vk_finalize_device(self._h, d_sum_ptr, d_rgb_ptr, n_floats, spp))

    def intersect(self, rays, medium_xi=None, flags=0):
        """``vk_intersect``: rays is a numpy array of RAY_DTYPE; returns HIT_DTYPE array."""
        rays = np.ascontiguousarray(rays, dtype=_abi.RAY_DTYPE)
        out = np.zeros(len(rays), dtype=_abi.HIT_DTYPE)
        xi = None
        if medium_xi is not None:
            xi = np.ascontiguousarray(medium_xi, dtype=np.float32)
            assert xi.size == len(rays) * _abi.VK_MEDIUM_XI_SLOTS
        self._check(self._L.vk_intersect(self._h, rays.ctypes.data, len(rays), xi.ctypes.data if xi is not None else None,
                                         flags, out.ctypes.data))
        return out

    def eval_batch(self, recs, flags=0):
        """``vk_eval_batch``: recs is a numpy array of EVAL_DTYPE (inputs filled in); returns a copy with the results."""
        out = np.ascontiguousarray(recs, dtype=_abi.EVAL_DTYPE).copy()
        self._check(self._L.vk_eval_batch(self._h, out.ctypes.data, len(out), flags))
        return out

    def measure_peaks(self):
        a, b = C.c_float(), C.c_float()
        self._check(self._L.vk_measure_peaks(self._h, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    def device_info(self):
        sm, khz = C.c_int(), C.c_int()
        name = C.create_string_buffer(128)
        self._check(self._L.vk_device_info(self._h, C.byref(sm), C.byref(khz), name, 128))
        return {"sm_count": sm.value, "clock_khz": khz.value, "name": name.value.decode()}

    def close(self):
        if getattr(self, "_h", None):
            self._L.vk_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiContext:
    """``vk_multi_create``: one context over several GPUs of the node (spp slices per device, the peers' integer
    accumulators added on the first device over NVLink).  Same frame as one GPU, bit for bit."""

    def __init__(self, devices):
        L = gpu_lib()
        h = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        rc = L.vk_multi_create(arr, len(devices), C.byref(h))
        if rc != 0:
            raise VecchioError(rc, (L.vk_multi_last_error(None) or b"").decode())
        self._h, self._L, self.devices = h, L, list(devices)

    def _check(self, rc):
        if rc != 0:
            raise VecchioError(rc, (self._L.vk_multi_last_error(self._h) or b"").decode())

    def upload(self, scene):
        self._check(self._L.vk_multi_scene_upload(self._h, scene.desc_ptr))

    def render(self, cam, params, want_sumsq=False):
        n = params.width * params.height * 3
        rgb = np.empty(n, dtype=np.float32)
        sq = np.empty(n, dtype=np.float32) if want_sumsq else None
        st = _abi.vk_stats()
        self._check(self._L.vk_multi_render(self._h, C.byref(cam), C.byref(params), rgb.ctypes.data,
                                            sq.ctypes.data if want_sumsq else None, C.byref(st)))
        shape = (params.height, params.width, 3)
        return rgb.reshape(shape), (sq.reshape(shape) if want_sumsq else None), st

    def render_rgb8(self, cam, params):
        out = np.empty(params.width * params.height * 3, dtype=np.uint8)
        st = _abi.vk_stats()
        self._check(self._L.vk_multi_render_rgb8(self._h, C.byref(cam), C.byref(params), out.ctypes.data, C.byref(st)))
        return out.reshape(params.height, params.width, 3), st

    def close(self):
        if getattr(self, "_h", None):
            self._L.vk_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def to_color(rgb):
    """``Vec3::to_color`` (src/vec3.rs:54-61) on an (..., 3) float array -> uint8 (output only)."""
    with np.errstate(invalid="ignore"):
        s = np.sqrt(rgb.astype(np.float32))
        s = np.where(s < 0.0, 0.0, np.where(s > 0.999, 0.999, s))  # Vec3::clamp keeps NaN
        v = np.float32(256.0) * s.astype(np.float32)
        v = np.where(np.isnan(v), 0.0, v)  # `NaN as u32` == 0
    return v.astype(np.uint32).astype(np.uint8)
