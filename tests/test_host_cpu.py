"""CPU-only tests (`-m "not gpu"`): the C ABI library loads and exports every symbol
include/b200rt.h declares, struct layouts match, host logic (Camera::init restatement, BVH
builder invariants, sample split) is right, and compute calls FAIL LOUDLY without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, SMALL_SCENES


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200rt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200rt_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from cpp_raytracer_b200 import capi
    lib = capi.lib()
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"libb200rt.so does not export {n}"
    assert sorted(capi.EXPORTS) == names, "capi.EXPORTS and include/b200rt.h disagree"
    assert lib.b200rt_version() == 2      # B200RT_VERSION: v2 added the multi-device entries and the Stats tail
    src = open(os.path.join(ROOT, "include", "b200rt.h")).read()
    assert "#define B200RT_VERSION 2" in src


def test_struct_layouts_match_header():
    from cpp_raytracer_b200 import capi
    assert capi.MATERIAL_DTYPE.itemsize == 40
    assert capi.SPHERE_DTYPE.itemsize == 40
    assert capi.QUAD_DTYPE.itemsize == 80
    assert capi.CAMERA_DTYPE.itemsize == 4 * 8 + 9 * 8 + 4 * 8 + 3 * 8 + 15 * 8
    assert ctypes.sizeof(capi.SceneDesc) == 48 and ctypes.sizeof(capi.RenderOpts) == 32


@pytest.mark.parametrize("name", SMALL_SCENES)
def test_camera_init_bit_exact_vs_reference(golden, name):
    """b200rt_camera_init restates Camera::init (camera.h:87-157); the golden scene files carry
    what the reference's init() derived for the same setter-level inputs."""
    import cpp_raytracer_b200 as rt
    scene = golden.scene(name)
    cam = scene.camera.copy()
    want = {k: cam[k].copy() for k in ("pixel00", "delta_x", "delta_y", "disk_x", "disk_y")}
    for k in want:
        cam[k] = 0
    got = rt.camera_init(cam)
    for k, v in want.items():
        assert np.array_equal(got[k], v), f"{name}: {k} differs from the reference's Camera::init"


def test_camera_init_rejects_bad_input(golden):
    import cpp_raytracer_b200 as rt
    cam = golden.scene("quads").camera.copy()
    cam["hfov"] = 1.0   # both FOVs given
    with pytest.raises(rt.B200rtError):
        rt.camera_init(cam)
    cam = golden.scene("quads").camera.copy()
    cam["image_w"] = 0
    with pytest.raises(rt.B200rtError):
        rt.camera_init(cam)


@pytest.mark.parametrize("name", SMALL_SCENES)
@pytest.mark.parametrize("leaf", [1, 2, 4, 8])
def test_bvh_builder_invariants(golden, name, leaf):
    """Every primitive in exactly one leaf, boxes nested and conservative (FP32 rounded outward),
    leaves type-pure, reported depth exact (so the traversal stack cannot overflow)."""
    import cpp_raytracer_b200 as rt
    scene = golden.scene(name)
    info = rt.selftest_bvh(scene, max_leaf_prims=leaf)
    assert info["n_prims"] == scene.n_prims
    assert info["stack_entries"] == 3 * info["tree_depth"] <= 128
    assert info["n_nodes"] >= 1


def test_bvh_builder_parallel_path_and_degenerate_inputs():
    """> 65536 primitives takes the thread-pool path; coincident centroids, zero-size and huge
    boxes must not break the builder."""
    import cpp_raytracer_b200 as rt
    from cpp_raytracer_b200 import capi
    rng = np.random.default_rng(5)
    n = 150_000
    sph = np.zeros(n, capi.SPHERE_DTYPE)
    sph["c"] = rng.uniform(-500, 500, (n, 3))
    sph["r"] = rng.uniform(0.05, 0.5, n)
    sph["c"][:1000] = [1.0, 2.0, 3.0]          # 1000 coincident spheres
    sph["r"][:1000] = 0.25
    sph["c"][1000] = [0, -1e6, 0]; sph["r"][1000] = 1e6
    sph["prim"] = np.arange(n)
    quads = np.zeros(64, capi.QUAD_DTYPE)
    quads["v"] = rng.uniform(-10, 10, (64, 3)); quads["s1"] = [1, 0, 0]; quads["s2"] = [0, 0, 1]
    quads["prim"] = n + np.arange(64)
    mats = np.zeros(1, capi.MATERIAL_DTYPE)
    cam = np.zeros(1, capi.CAMERA_DTYPE)
    scene = capi.HostScene(mats, sph, quads, cam)
    for threads in (1, 4):
        info = rt.selftest_bvh(scene, threads=threads)
        assert info["n_prims"] == n + 64 and info["tree_depth"] <= 40


def test_empty_scene_builds():
    import cpp_raytracer_b200 as rt
    from cpp_raytracer_b200 import capi
    scene = capi.HostScene(np.zeros(0, capi.MATERIAL_DTYPE), np.zeros(0, capi.SPHERE_DTYPE),
                           np.zeros(0, capi.QUAD_DTYPE), np.zeros(1, capi.CAMERA_DTYPE))
    info = rt.selftest_bvh(scene)
    assert info["n_nodes"] == 1 and info["tree_depth"] == 1


def test_bad_scene_is_rejected(golden):
    import cpp_raytracer_b200 as rt
    from cpp_raytracer_b200 import capi
    scene = golden.scene("cornell")
    bad = capi.HostScene(scene.materials.copy(), scene.spheres, scene.quads.copy(), scene.camera)
    bad.materials["kind"][1] = 9
    with pytest.raises(rt.B200rtError) as e:
        rt.selftest_bvh(bad)
    assert e.value.code == capi.EINVAL and "unknown material kind" in str(e.value)


def test_compute_fails_loudly_without_gpu(golden):
    """No CPU fallback: without a device, scene creation / tone map return ENODEVICE."""
    import cpp_raytracer_b200 as rt
    from cpp_raytracer_b200 import capi
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(rt.B200rtError) as e:
        rt.DeviceSceneHandle(golden.scene("quads"))
    assert e.value.code == capi.ENODEVICE
    with pytest.raises(rt.B200rtError) as e:
        rt.tonemap(np.zeros((4, 3), np.float32))
    assert e.value.code == capi.ENODEVICE
    # the entries added in ABI v2 behave the same: the multi-device forms, the one-call render, the test hooks
    scene = golden.scene("quads")
    cam = rt.camera_with(scene.camera, image_w=8, image_h=8, spp=1)
    with pytest.raises(rt.B200rtError) as e:
        rt.DeviceSceneHandle(scene, devices=[0, 1])
    assert e.value.code == capi.ENODEVICE
    for devices in (None, [0], [0, 0]):
        with pytest.raises(rt.B200rtError) as e:
            rt.render_scene(scene, cam, devices=devices)
        assert e.value.code == capi.ENODEVICE
    with pytest.raises(rt.B200rtError) as e:
        capi.debug_philox(np.zeros((1, 6), np.uint32))
    assert e.value.code == capi.ENODEVICE
    with pytest.raises(rt.B200rtError) as e:
        capi.debug_samplers(np.zeros((1, 2), np.uint32))
    assert e.value.code == capi.ENODEVICE
    assert capi.lib().b200rt_trim() == 0          # nothing cached, nothing to do: not an error


def test_scene_description_is_validated_before_any_device_is_needed(golden):
    """Malformed descriptions are EINVAL whether or not a GPU is present (structure checks come first)."""
    import ctypes
    import cpp_raytracer_b200 as rt
    from cpp_raytracer_b200 import capi
    h = ctypes.c_void_p()
    assert capi.lib().b200rt_scene_create(None, None, ctypes.byref(h)) == capi.EINVAL
    desc = capi.SceneDesc(0, 3, 0, None, None, None)        # a count without an array
    assert capi.lib().b200rt_scene_create(ctypes.byref(desc), None, ctypes.byref(h)) == capi.EINVAL
    assert capi.lib().b200rt_scene_create_multi(ctypes.byref(desc), None, None, 2, ctypes.byref(h)) == capi.EINVAL
    assert b"count without an array" in capi.lib().b200rt_last_error()


def test_sample_ranges_tile_exactly():
    from cpp_raytracer_b200.dist import sample_range
    for spp in (1, 7, 1024, 1000, 4096):
        for world in (1, 2, 3, 4, 8):
            got = [sample_range(spp, r, world) for r in range(world)]
            assert got[0][0] == 0 and sum(c for _, c in got) == spp
            for (a, ca), (b, _) in zip(got, got[1:]):
                assert a + ca == b


def test_scene_file_roundtrip(golden, tmp_path):
    from cpp_raytracer_b200 import scene_io
    s = golden.scene("xmas")
    p = str(tmp_path / "x.scene.gz")
    scene_io.save_scene(s, p)
    t = scene_io.load_scene(p)
    assert np.array_equal(s.spheres, t.spheres) and np.array_equal(s.quads, t.quads)
    assert np.array_equal(s.materials, t.materials) and s.camera.tobytes() == t.camera.tobytes()


def test_header_is_plain_c_and_bindings_match_its_layout(tmp_path):
    """include/b200rt.h compiled as C (gcc -std=c11 -pedantic): sizes and selected field offsets of every
    struct, as the compiler lays them out, against the ctypes / numpy mirrors in capi.py."""
    import subprocess
    from cpp_raytracer_b200 import capi
    structs = ["B200rtMaterial", "B200rtSphere", "B200rtQuad", "B200rtCamera", "B200rtSceneDesc", "B200rtBuildOpts",
               "B200rtSceneInfo", "B200rtRenderOpts", "B200rtStats", "B200rtShadeRecord"]
    offsets = [("B200rtShadeRecord", "t"), ("B200rtShadeRecord", "atten"), ("B200rtShadeRecord", "emit"),
               ("B200rtShadeRecord", "prim"), ("B200rtShadeRecord", "flags"), ("B200rtCamera", "pixel00"),
               ("B200rtCamera", "background"), ("B200rtSphere", "mat"), ("B200rtQuad", "mat"), ("B200rtMaterial", "param"),
               ("B200rtStats", "paths"), ("B200rtSceneInfo", "build_ms"), ("B200rtRenderOpts", "flags")]
    src = tmp_path / "layout.c"
    body = "".join(f'    printf("sizeof {s} %zu\\n", sizeof({s}));\n' for s in structs)
    body += "".join(f'    printf("offsetof {s}.{f} %zu\\n", offsetof({s}, {f}));\n' for s, f in offsets)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "b200rt.h"\nint main(void) {\n' + body + "    return 0;\n}\n")
    exe = str(tmp_path / "layout")
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c11", "-pedantic", "-Wall", "-Werror", f"-I{inc}", "-o", exe, str(src)], check=True, env=env)
    got = {}
    for line in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines():
        kind, name, val = line.split()
        got[(kind, name)] = int(val)
    mirror = {"B200rtMaterial": capi.MATERIAL_DTYPE.itemsize, "B200rtSphere": capi.SPHERE_DTYPE.itemsize,
              "B200rtQuad": capi.QUAD_DTYPE.itemsize, "B200rtCamera": capi.CAMERA_DTYPE.itemsize,
              "B200rtSceneDesc": ctypes.sizeof(capi.SceneDesc), "B200rtBuildOpts": ctypes.sizeof(capi.BuildOpts),
              "B200rtSceneInfo": ctypes.sizeof(capi.SceneInfo), "B200rtRenderOpts": ctypes.sizeof(capi.RenderOpts),
              "B200rtStats": ctypes.sizeof(capi.Stats), "B200rtShadeRecord": capi.SHADE_RECORD_DTYPE.itemsize}
    for s, size in mirror.items():
        assert got[("sizeof", s)] == size, (s, got[("sizeof", s)], size)
    np_of = {"B200rtShadeRecord": capi.SHADE_RECORD_DTYPE, "B200rtCamera": capi.CAMERA_DTYPE, "B200rtSphere": capi.SPHERE_DTYPE,
             "B200rtQuad": capi.QUAD_DTYPE, "B200rtMaterial": capi.MATERIAL_DTYPE}
    ct_of = {"B200rtStats": capi.Stats, "B200rtSceneInfo": capi.SceneInfo, "B200rtRenderOpts": capi.RenderOpts}
    for s, f in offsets:
        want = np_of[s].fields[f][1] if s in np_of else getattr(ct_of[s], f).offset
        assert got[("offsetof", f"{s}.{f}")] == want, (s, f)
