"""ctypes wrapper of oracle/libpt_oracle.so (the plain-C restatement, pt_oracle.c).
TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg, never by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libpt_oracle.so")


class OScene(C.Structure):
    _fields_ = [("n_materials", C.c_uint64), ("n_spheres", C.c_uint64), ("n_quads", C.c_uint64),
                ("materials", C.c_void_p), ("spheres", C.c_void_p), ("quads", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "pt_oracle.c")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            env = dict(os.environ)
            env.pop("CC", None); env.pop("CXX", None)
            subprocess.run(["make", "-C", HERE, "port"], check=True, env=env, capture_output=True)
        l = C.CDLL(LIB)
        l.oracle_rand_double.restype = C.c_double
        l.oracle_rand_double.argtypes = [C.POINTER(C.c_uint32), C.c_double, C.c_double]
        l.oracle_seed_sequence_next.restype = C.c_uint32
        l.oracle_seed_sequence_next.argtypes = [C.POINTER(C.c_uint32)]
        l.oracle_camera_init.argtypes = [C.c_void_p]
        l.oracle_raycast_brute.argtypes = [C.POINTER(OScene), C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        l.oracle_prim_hit.restype = C.c_double
        l.oracle_prim_hit.argtypes = [C.POINTER(OScene), C.c_uint32, C.c_void_p, C.c_double, C.c_double]
        l.oracle_reflected.argtypes = [C.c_void_p] * 3
        l.oracle_refracted.restype = C.c_int
        l.oracle_refracted.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        l.oracle_reflectance.restype = C.c_double
        l.oracle_reflectance.argtypes = [C.c_double, C.c_double]
        l.oracle_render.restype = C.c_uint64
        l.oracle_render.argtypes = [C.POINTER(OScene), C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p]
        l.oracle_camera_ray.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p]
        l.oracle_camera_ray.restype = None
        l.oracle_tonemap.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        l.oracle_random_unit_vectors.argtypes = [C.POINTER(C.c_uint32), C.c_int64, C.c_void_p]
        l.oracle_random_vectors_in_unit_disk.argtypes = [C.POINTER(C.c_uint32), C.c_int64, C.c_void_p]
        _lib = l
    return _lib


def _oscene(scene):
    d = scene.desc()   # same packed struct layouts as the C ABI
    return OScene(d.n_materials, d.n_spheres, d.n_quads, d.materials, d.spheres, d.quads)


def rand_doubles(state: int, n: int, lo=0.0, hi=1.0):
    st = C.c_uint32(state)
    out = [lib().oracle_rand_double(C.byref(st), lo, hi) for _ in range(n)]
    return out, st.value


def seed_sequence_next(seed: int) -> int:
    st = C.c_uint32(seed)
    return lib().oracle_seed_sequence_next(C.byref(st))


def camera_init(cam: np.ndarray) -> np.ndarray:
    out = np.array(cam).reshape(1).copy()
    lib().oracle_camera_init(out.ctypes.data)
    return out


def raycast_brute(scene, rays, tmin=1e-5, tmax=float("inf")):
    rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
    n = len(rays)
    prim = np.empty(n, np.int32)
    t = np.empty(n, np.float64)
    osc = _oscene(scene)
    lib().oracle_raycast_brute(C.byref(osc), rays.ctypes.data, n, tmin, tmax, prim.ctypes.data, t.ctypes.data)
    return prim, t


def prim_hit(scene, prim: int, ray, tmin=1e-5, tmax=float("inf")) -> float:
    ray = np.ascontiguousarray(ray, dtype=np.float64)
    osc = _oscene(scene)
    return lib().oracle_prim_hit(C.byref(osc), prim, ray.ctypes.data, tmin, tmax)


def reflected(d, n):
    d = np.ascontiguousarray(d, np.float64); n = np.ascontiguousarray(n, np.float64)
    out = np.empty(3)
    lib().oracle_reflected(d.ctypes.data, n.ctypes.data, out.ctypes.data)
    return out


def refracted(d, n, ratio):
    d = np.ascontiguousarray(d, np.float64); n = np.ascontiguousarray(n, np.float64)
    out = np.empty(3)
    ok = lib().oracle_refracted(d.ctypes.data, n.ctypes.data, ratio, out.ctypes.data)
    return out if ok else None


def reflectance(c, ratio):
    return lib().oracle_reflectance(c, ratio)


def camera_ray(cam, row: int, col: int, vx: float, vy: float, r1: float, r2: float):
    """random_ray_through_pixel (camera.h:184-200) with the draws supplied; cam must be initialised."""
    cam = np.ascontiguousarray(cam).reshape(1)
    out = np.empty(6)
    lib().oracle_camera_ray(cam.ctypes.data, row, col, vx, vy, r1, r2, out.ctypes.data)
    return out


def render(scene, cam, lcg_state: int):
    cam = np.ascontiguousarray(cam).reshape(1)
    h, w = int(cam["image_h"][0]), int(cam["image_w"][0])
    out = np.empty((h, w, 3), np.float64)
    st = C.c_uint32(lcg_state)
    osc = _oscene(scene)
    rays = lib().oracle_render(C.byref(osc), cam.ctypes.data, C.byref(st), out.ctypes.data)
    return out, int(rays), st.value


def random_unit_vectors(n: int, lcg_state: int = 12345) -> np.ndarray:
    """n draws of Vec3D::random_unit_vector() (vec3d.h:64-75, rejection + normalise) from the reference's LCG."""
    out = np.empty((n, 3), np.float64)
    st = C.c_uint32(lcg_state)
    lib().oracle_random_unit_vectors(C.byref(st), n, out.ctypes.data)
    return out


def random_vectors_in_unit_disk(n: int, lcg_state: int = 12345) -> np.ndarray:
    """n draws of Vec3D::random_vector_in_unit_disk() (vec3d.h:79-85)."""
    out = np.empty((n, 2), np.float64)
    st = C.c_uint32(lcg_state)
    lib().oracle_random_vectors_in_unit_disk(C.byref(st), n, out.ctypes.data)
    return out


def tonemap(rgb):
    rgb = np.ascontiguousarray(rgb, np.float64)
    out = np.empty(rgb.shape, np.int32)
    lib().oracle_tonemap(rgb.ctypes.data, rgb.size // 3, out.ctypes.data)
    return out
