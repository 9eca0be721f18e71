"""Cull effectiveness and K1 timing with the FP32 miss test on/off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from terminalraytracer_b200 import renderer as R, scene as S
w = int(sys.argv[1]) if len(sys.argv) > 1 else 1920
h = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
kind = sys.argv[3] if len(sys.argv) > 3 else "demo"
sky = S.get_skybox("milky_way")
rd = R.Renderer(0, sky)
sc = S.SceneData(w, h, sky, kind=kind).set_time(3.7)
rd.set_scene(sc)
ctr, F = rd.count_rows(w, h, 0, h)
print("sphere tests %d exact %d (%.1f%%) disc_ok %d (%.1f%%) violations %d; F/sample %.1f" % (
    ctr[0], ctr[27], 100.0 * ctr[27] / max(ctr[0], 1), ctr[1], 100.0 * ctr[1] / max(ctr[0], 1), ctr[28], F / (10.0 * w * h)))
for cull in (1, 0, 1):
    rd.L.trt_set_cull(cull)
    for i in range(3):
        rd.render_ansi(sc)
    ms = rd.last_ms()[0]
    print("cull=%d K1 ms %.3f  Mrays/s %.0f  model TFLOP/s %.2f" % (cull, ms, 10 * w * h / ms / 1e3, F / ms / 1e9))
rd.close()
