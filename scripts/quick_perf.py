"""Development aid: K1 time of the demo scene at three sizes and of the 1024-sphere stress scene (one GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from terminalraytracer_b200 import renderer as R, scene as S
sky = S.get_skybox("milky_way")
rd = R.Renderer(0, sky)
CASES = [(1920, 1080, "demo", 3), (7680, 4320, "demo", 3)] if os.environ.get("QUICK") else [(1920, 1080, "demo", 3), (3840, 2160, "demo", 3), (7680, 4320, "demo", 3), (480, 270, "stress", 3), (1920, 1080, "stress", 1)]
for (w, h, kind, n) in CASES:
    sc = S.SceneData(w, h, sky, kind=kind).set_time(3.7)
    best = 1e30
    for i in range(n):
        rd.render_ansi(sc)
        best = min(best, rd.last_ms()[0])
    print("%-6s %5dx%-5d K1 %9.3f ms  %8.1f Mrays/s" % (kind, w, h, best, 10 * w * h / best / 1e3), flush=True)
rd.close()
