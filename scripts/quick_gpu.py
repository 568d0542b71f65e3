import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpp_raytracer_b200 as rt
from cpp_raytracer_b200 import scene_io
import os
G=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),'tests','golden')
for name in ["rtow_final", "rtow_lights", "quads", "cornell_empty", "cornell", "xmas", "pathological"]:
    s = scene_io.load_scene(f'{G}/{name}.scene.gz')
    rays,tmin,tmax = scene_io.load_rays(f'{G}/{name}.rays.gz')
    pb,tb = scene_io.load_hits(f'{G}/{name}.hits_brute.gz')
    with rt.DeviceSceneHandle(s) as d:
        print(name, d.info())
        p,t = d.raycast(rays,tmin,tmax)
        bad = np.nonzero((p!=pb)|(t!=tb))[0]
        print('  raycast mismatches vs brute:', len(bad), 'of', len(rays))
        for i in bad[:5]: print('   ', i, rays[i], p[i], repr(t[i]), pb[i], repr(tb[i]))
        if name != 'pathological':
            cam = rt.camera_with(s.camera, image_w=160, image_h=90 if 'rtow' in name or name=='xmas' else 160, spp=64)
            img, st = d.render(cam)
            print('  render', st, 'mean', img.mean(axis=(0,1)))
