"""ctypes binding of the CPU oracle (oracle/liboracle_rt.so).  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs import this module; the product never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle_rt.so")


class Camera(C.Structure):
    _fields_ = [("eye", C.c_float * 3), ("right", C.c_float * 3), ("up", C.c_float * 3), ("forward", C.c_float * 3)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("primary_rays", "shadow_rays", "primary_hits", "bound_tests",
                                          "leaf_tests", "disc_nonneg", "hit_updates")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}

    def flop_per_ray(self):
        """17 T + 3 P + 19 U + 20 f_prim + 30 f_shadow per ray."""
        rays = self.primary_rays + self.shadow_rays
        if rays == 0:
            return 0.0
        tests = self.bound_tests + self.leaf_tests
        return (17.0 * tests + 3.0 * self.disc_nonneg + 19.0 * self.hit_updates
                + 20.0 * self.primary_rays + 30.0 * self.shadow_rays) / rays


class Ray(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("dir", C.c_float * 3)]


class Hit(C.Structure):
    _fields_ = [("distance", C.c_float), ("normal", C.c_float * 3)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "liboracle_rt.so"])


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    L = C.CDLL(LIB_PATH)
    vp, u32, f32p, u8p = C.c_void_p, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_uint8)
    L.orc_scene_create.restype = vp
    L.orc_scene_create.argtypes = [u32, f32p, C.c_float, f32p, f32p]
    L.orc_scene_create_default.restype = vp
    L.orc_scene_create_default.argtypes = []
    L.orc_scene_create_from_nodes.restype = vp
    L.orc_scene_create_from_nodes.argtypes = [u32, f32p, C.POINTER(u32), f32p, f32p]
    L.orc_scene_destroy.argtypes = [vp]
    L.orc_scene_counts.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.orc_scene_flatten.restype = u32
    L.orc_scene_flatten.argtypes = [vp, f32p, C.POINTER(u32), u32]
    L.orc_scene_light.argtypes = [vp, f32p]
    L.orc_scene_eye.argtypes = [vp, f32p]
    L.orc_sphere_distance_from_ray.restype = C.c_float
    L.orc_sphere_distance_from_ray.argtypes = [f32p, C.c_float, C.POINTER(Ray)]
    L.orc_sphere_intersect.argtypes = [f32p, C.c_float, C.POINTER(Hit), C.POINTER(Ray)]
    L.orc_vec_normalized.argtypes = [f32p, f32p]
    L.orc_vec_len.restype = C.c_float
    L.orc_vec_len.argtypes = [f32p]
    L.orc_trace_rays.argtypes = [vp, C.c_size_t, C.POINTER(Ray), C.POINTER(Hit)]
    L.orc_render_region.argtypes = [vp, C.POINTER(Camera), u32, u32, u32, u32, u32, u32, u32, u8p, u8p,
                                    C.POINTER(Counters)]
    L.orc_render.argtypes = [vp, C.POINTER(Camera), u32, u32, u32, u32, u8p, C.POINTER(Counters)]
    L.orc_render_rows.argtypes = [vp, C.POINTER(Camera), u32, u32, u32, u32, u32, u32, u32, u8p,
                                  C.POINTER(Counters)]
    _lib = L
    return L


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def make_camera(eye, right=(1, 0, 0), up=(0, 1, 0), forward=(0, 0, 1)):
    cam = Camera()
    cam.eye[:] = [float(x) for x in eye]
    cam.right[:] = [float(x) for x in right]
    cam.up[:] = [float(x) for x in up]
    cam.forward[:] = [float(x) for x in forward]
    return cam


class Scene:
    """Scene (render.rs:138-167) held by the oracle."""

    def __init__(self, level=8, origin=(0.0, -1.0, 0.0), radius=1.0, light=(-1.0, -3.0, 2.0), eye=(0.0, 0.0, -4.0),
                 _handle=None):
        L = lib()
        self._h = _handle if _handle is not None else L.orc_scene_create(level, _f3(origin), radius, _f3(light), _f3(eye))
        if not self._h:
            raise ValueError("oracle scene creation failed (level must be > 1)")

    @classmethod
    def from_nodes(cls, spheres4, skip, light, eye):
        sph = np.ascontiguousarray(spheres4, dtype=np.float32)
        sk = np.ascontiguousarray(skip, dtype=np.uint32)
        h = lib().orc_scene_create_from_nodes(len(sk), _fp(sph), sk.ctypes.data_as(C.POINTER(C.c_uint32)),
                                              _f3(light), _f3(eye))
        return cls(_handle=h)

    def __del__(self):
        try:
            if self._h:
                lib().orc_scene_destroy(self._h)
        except Exception:
            pass
        self._h = None

    def counts(self):
        g, i = C.c_uint64(), C.c_uint64()
        lib().orc_scene_counts(self._h, C.byref(g), C.byref(i))
        return g.value, i.value

    def flatten(self):
        n = lib().orc_scene_flatten(self._h, None, None, 0)
        sph = np.empty((n, 4), np.float32)
        sk = np.empty(n, np.uint32)
        lib().orc_scene_flatten(self._h, _fp(sph), sk.ctypes.data_as(C.POINTER(C.c_uint32)), n)
        return sph, sk

    def light(self):
        v = (C.c_float * 3)()
        lib().orc_scene_light(self._h, v)
        return np.array(v[:], np.float32)

    def eye(self):
        v = (C.c_float * 3)()
        lib().orc_scene_eye(self._h, v)
        return np.array(v[:], np.float32)

    def trace_rays(self, pos, dirs):
        pos = np.asarray(pos, np.float32).reshape(-1, 3)
        dirs = np.asarray(dirs, np.float32).reshape(-1, 3)
        n = len(pos)
        rays = np.ascontiguousarray(np.concatenate([pos, dirs], axis=1), np.float32)
        hits = np.empty((n, 4), np.float32)
        lib().orc_trace_rays(self._h, n, rays.ctypes.data_as(C.POINTER(Ray)), hits.ctypes.data_as(C.POINTER(Hit)))
        return hits[:, 0].copy(), hits[:, 1:].copy()

    def render_region(self, width, height, spp, l, b, r, t, camera=None, kinds=False):
        out = np.empty(((t - b), (r - l), 4), np.uint8)
        kd = np.empty(((t - b), (r - l), spp * spp), np.uint8) if kinds else None
        ctr = Counters()
        lib().orc_render_region(self._h, C.byref(camera) if camera is not None else None, width, height, spp,
                                l, b, r, t, _u8p(out), _u8p(kd) if kinds else None, C.byref(ctr))
        return (out, kd, ctr) if kinds else (out, ctr)

    def render(self, width, height, spp, threads=None, camera=None):
        threads = threads or os.cpu_count() or 1
        out = np.empty((height, width, 4), np.uint8)
        ctr = Counters()
        lib().orc_render(self._h, C.byref(camera) if camera is not None else None, width, height, spp, threads,
                         _u8p(out), C.byref(ctr))
        return out, ctr

    def render_rows(self, width, height, spp, row_start, row_stride, row_count, threads=None, camera=None):
        threads = threads or os.cpu_count() or 1
        out = np.empty((row_count, width, 4), np.uint8)
        ctr = Counters()
        lib().orc_render_rows(self._h, C.byref(camera) if camera is not None else None, width, height, spp,
                              row_start, row_stride, row_count, threads, _u8p(out), C.byref(ctr))
        return out, ctr


def sphere_distance_from_ray(center, radius, pos, direction):
    r = Ray()
    r.pos[:] = [float(x) for x in pos]
    r.dir[:] = [float(x) for x in direction]
    return float(lib().orc_sphere_distance_from_ray(_f3(center), radius, C.byref(r)))


def sphere_intersect(center, radius, hit_distance, pos, direction, hit_normal=(0, 0, 0)):
    r = Ray()
    r.pos[:] = [float(x) for x in pos]
    r.dir[:] = [float(x) for x in direction]
    h = Hit()
    h.distance = hit_distance
    h.normal[:] = [float(x) for x in hit_normal]
    lib().orc_sphere_intersect(_f3(center), radius, C.byref(h), C.byref(r))
    return float(h.distance), tuple(h.normal[:])


def vec_normalized(v):
    out = (C.c_float * 3)()
    lib().orc_vec_normalized(_f3(v), out)
    return tuple(out[:])


def vec_len(v):
    return float(lib().orc_vec_len(_f3(v)))
