"""Development aid for compute-sanitizer: a few small frames through every K1 flavour and K2 (demo, stress, cull off, orbit sink)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from terminalraytracer_b200 import renderer as R, scene as S
sky = S.synthetic_cubemap("uv_gradient", 64)
rd = R.Renderer(0, sky)
for sc in (S.SceneData(97, 53, sky).set_time(3.7), S.SceneData(50, 30, sky, kind="stress", num_spheres=200).set_time(3.7)):
    a = rd.project_scene(sc)
    b = np.array(rd.render_ansi(sc))
    rd.set_scene(sc)
    ctr, _ = rd.count_rows(sc.width, sc.height, 0, sc.height)
    rd.L.trt_set_cull(0)
    c = rd.project_scene(sc)
    rd.L.trt_set_cull(1)
    print(sc.width, sc.height, "same with cull off:", np.array_equal(a, c), "violations", ctr[28], "bytes", b.size)
got = []
rd.render_orbit(S.SceneData(64, 36, sky), [0.0, 1.0, 2.0], lambda f, v: got.append(int(v.sum())) and False)
print("orbit", got)
rd.close()
