"""Experiment: traversal work and kernel time vs leaf size (run on the GPU box)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import cpp_raytracer_b200 as rt
from cpp_raytracer_b200 import scene_io, capi
G = os.path.join(ROOT, 'tests', 'golden')
for name, w, h, spp in [('rtow_lights', 960, 540, 64), ('xmas', 960, 540, 64), ('cornell', 512, 512, 64)]:
    s = scene_io.load_scene(f'{G}/{name}.scene.gz')
    cam = rt.camera_with(s.camera, image_w=w, image_h=h, spp=spp)
    for leaf in (1, 2, 3, 4, 6, 8):
        with rt.DeviceSceneHandle(s, max_leaf_prims=leaf) as d:
            info = d.info()
            d.render(cam)
            _, st = d.render(cam)
            _, sc = d.render(cam, flags=capi.FLAG_COUNTERS)
            print(json.dumps({'scene': name, 'leaf': leaf, 'nodes': info['n_nodes'], 'depth': info['tree_depth'],
                              'kernel_ms': round(st['kernel_ms'], 3), 'mpaths_s': round(st['paths'] / st['kernel_ms'] / 1e3, 1),
                              'rays_per_path': round(st['rays'] / st['paths'], 3),
                              'nodes_per_ray': round(sc['node_visits'] / sc['rays'], 2),
                              'prims_per_ray': round(sc['prim_tests'] / sc['rays'], 2)}), flush=True)
