"""Small driver for ncu: a few fused frames (K1 + K2) of the demo scene.  usage: profile_k1.py [W H [frames]]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from terminalraytracer_b200 import renderer as R, scene as S
w = int(sys.argv[1]) if len(sys.argv) > 1 else 1920
h = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
sky = S.get_skybox("milky_way")
rd = R.Renderer(0, sky)
sc = S.SceneData(w, h, sky).set_time(3.7)
for i in range(n):
    rd.render_ansi(sc)
    print("frame", i, "K1 ms %.3f K2 ms %.3f" % rd.last_ms(), "Mrays/s %.0f" % (10 * w * h / rd.last_ms()[0] / 1e3))
rd.close()
