"""Development aid: per-frame timeline of OrbitPipeline.stream over the whole 360-frame orbit (one GPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from terminalraytracer_b200 import pipeline, renderer as R, scene as S, sharding
w, h, n = 1920, 1080, 360
sky = S.get_skybox("milky_way")
rd = R.Renderer(0, sky)
times = sharding.orbit_times(360)
orbit = pipeline.OrbitPipeline(rd, w, h)
orbit.stream(S.SceneData(w, h, sky), times[:8], None)
devnull = os.open(os.devnull, os.O_WRONLY)
for rep in range(2):
    stamps = []
    def write(k, v):
        stamps.append(time.perf_counter())
        return os.write(devnull, v) != len(v)
    t0 = time.perf_counter()
    orbit.stream(S.SceneData(w, h, sky), times, write)
    t1 = time.perf_counter()
    s = np.array(stamps) - t0
    print("rep %d: %.2f ms/frame overall; per 30 frames (ms/frame): %s" % (rep, (t1 - t0) * 1e3 / n, " ".join("%.2f" % ((s[i + 29] - s[i]) * 1e3 / 29) for i in range(0, n - 29, 30))), flush=True)
# the C loop alone over the same path (library-owned pinned buffers, no-op python sink)
stamps = []
t0 = time.perf_counter()
rd.render_orbit(S.SceneData(w, h, sky), times, lambda f, v: stamps.append(time.perf_counter()) and False)
s = np.array(stamps) - t0
print("trt_render_orbit: %.2f ms/frame overall; per 30 frames: %s" % ((time.perf_counter() - t0) * 1e3 / n, " ".join("%.2f" % ((s[i + 29] - s[i]) * 1e3 / 29) for i in range(0, n - 29, 30))), flush=True)
orbit.close()
rd.close()
