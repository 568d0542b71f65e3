"""Flat scene files (.scene) and ray files shared by the product, the tests and the oracle bridge.

.scene layout (little endian; written by oracle/ref_bridge.cpp `dump` from reference-built
objects, and by save_scene() here):
    8  bytes  magic "B2RTSCN1"
    3  u64    n_materials, n_spheres, n_quads
    1  CAMERA_DTYPE record          (setter-level fields + the fields Camera::init derives)
    n_materials x MATERIAL_DTYPE, n_spheres x SPHERE_DTYPE, n_quads x QUAD_DTYPE
The three array dtypes are exactly the C ABI structs of include/b200rt.h.

ray file: u64 n, f64 tmin, f64 tmax, n x 6 f64 (origin, direction).
hit file (oracle output): n x i32 prim, n x f64 t.
"""
from __future__ import annotations

import gzip
import os

import numpy as np

from .capi import CAMERA_DTYPE, MATERIAL_DTYPE, QUAD_DTYPE, SPHERE_DTYPE, HostScene

MAGIC = b"B2RTSCN1"


def _open(path, mode):
    return gzip.open(path, mode) if str(path).endswith(".gz") else open(path, mode)


def load_scene(path: str) -> HostScene:
    with _open(path, "rb") as f:
        buf = f.read()
    if buf[:8] != MAGIC:
        raise ValueError(f"{path}: not a B2RTSCN1 scene file")
    n_mat, n_sph, n_quad = np.frombuffer(buf, dtype="<u8", count=3, offset=8)
    off = 32
    cam = np.frombuffer(buf, dtype=CAMERA_DTYPE, count=1, offset=off).copy()
    off += CAMERA_DTYPE.itemsize
    mats = np.frombuffer(buf, dtype=MATERIAL_DTYPE, count=int(n_mat), offset=off).copy()
    off += MATERIAL_DTYPE.itemsize * int(n_mat)
    sph = np.frombuffer(buf, dtype=SPHERE_DTYPE, count=int(n_sph), offset=off).copy()
    off += SPHERE_DTYPE.itemsize * int(n_sph)
    quads = np.frombuffer(buf, dtype=QUAD_DTYPE, count=int(n_quad), offset=off).copy()
    name = os.path.basename(str(path)).split(".")[0]
    return HostScene(mats, sph, quads, cam, name)


def save_scene(scene: HostScene, path: str) -> None:
    with _open(path, "wb") as f:
        f.write(MAGIC)
        f.write(np.array([len(scene.materials), len(scene.spheres), len(scene.quads)], dtype="<u8").tobytes())
        f.write(np.ascontiguousarray(scene.camera, dtype=CAMERA_DTYPE).reshape(1).tobytes())
        f.write(np.ascontiguousarray(scene.materials, dtype=MATERIAL_DTYPE).tobytes())
        f.write(np.ascontiguousarray(scene.spheres, dtype=SPHERE_DTYPE).tobytes())
        f.write(np.ascontiguousarray(scene.quads, dtype=QUAD_DTYPE).tobytes())


def save_rays(path: str, rays: np.ndarray, tmin: float = 1e-5, tmax: float = float("inf")) -> None:
    rays = np.ascontiguousarray(rays, dtype="<f8").reshape(-1, 6)
    with _open(path, "wb") as f:
        f.write(np.array([rays.shape[0]], dtype="<u8").tobytes())
        f.write(np.array([tmin, tmax], dtype="<f8").tobytes())
        f.write(rays.tobytes())


def load_rays(path: str):
    with _open(path, "rb") as f:
        buf = f.read()
    n = int(np.frombuffer(buf, dtype="<u8", count=1)[0])
    tmin, tmax = np.frombuffer(buf, dtype="<f8", count=2, offset=8)
    rays = np.frombuffer(buf, dtype="<f8", count=n * 6, offset=24).reshape(n, 6).copy()
    return rays, float(tmin), float(tmax)


def load_hits(path: str):
    with _open(path, "rb") as f:
        buf = f.read()
    n = len(buf) // 12
    prim = np.frombuffer(buf, dtype="<i4", count=n).copy()
    t = np.frombuffer(buf, dtype="<f8", count=n, offset=4 * n).copy()
    return prim, t


def load_hdr(path: str) -> np.ndarray:
    """Reference render written by the bridge: u64 w, u64 h, then h*w*3 float32 (or float64)."""
    with _open(path, "rb") as f:
        buf = f.read()
    w, h = (int(x) for x in np.frombuffer(buf, dtype="<u8", count=2))
    n = w * h * 3
    if len(buf) - 16 == n * 8:
        return np.frombuffer(buf, dtype="<f8", count=n, offset=16).reshape(h, w, 3).copy()
    return np.frombuffer(buf, dtype="<f4", count=n, offset=16).reshape(h, w, 3).copy()


def camera_rays(cam: np.ndarray, n: int, seed: int = 0) -> np.ndarray:
    """n primary rays of `cam` with a fixed jitter table (pinhole origin), as doubles:
    random_ray_through_pixel (reference camera.h:184-200) with the randoms drawn here."""
    cam = np.asarray(cam).reshape(1)[0]
    rng = np.random.default_rng(seed)
    w, h = int(cam["image_w"]), int(cam["image_h"])
    col = rng.integers(0, w, n).astype(np.float64)
    row = rng.integers(0, h, n).astype(np.float64)
    jx = rng.random(n) - 0.5
    jy = rng.random(n) - 0.5
    p = (cam["pixel00"][None, :] + row[:, None] * cam["delta_y"][None, :] + col[:, None] * cam["delta_x"][None, :]
         + jx[:, None] * cam["delta_x"][None, :] + jy[:, None] * cam["delta_y"][None, :])
    o = np.broadcast_to(cam["center"][None, :], p.shape)
    return np.concatenate([o, p - o], axis=1)
