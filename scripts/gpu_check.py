"""Quick on-box sanity: GPU path vs oracle/_ref on small frames (development aid, not a test)."""
import ctypes as C, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from terminalraytracer_b200 import abi, scene as S, renderer as R

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
orc = C.CDLL(os.path.join(root, "oracle/_build/libtrt_oracle.so"))
ref = C.CDLL(os.path.join(root, "oracle/_ref/libtrt_ref.so"))

def cpu(lib, fn, sc):
    px = np.zeros((sc.height, sc.width, 3))
    scr = abi.Screen(px.ctypes.data_as(C.POINTER(abi.Vector)), sc.width, sc.height)
    t0 = time.time(); getattr(lib, fn)(C.byref(sc.c), C.byref(scr)); return px, time.time() - t0

rd = R.Renderer(0)
print("peaks", rd.measure_peaks())
for name in ["uv_checker", "colors", "milky_way"]:
    sky = S.get_skybox(name)
    rd.upload_skybox(sky)
    for (w, h, t) in [(96, 56, 3.7), (480, 280, 0.0), (480, 280, 3.7)]:
        sc = S.SceneData(w, h, sky).set_time(t)
        g = rd.project_scene(sc)
        ms = rd.last_ms()[0]
        o, to = cpu(orc, "orc_project_scene", sc)
        r, tr = cpu(ref, "project_scene", sc)
        print(name, w, h, t, "gpu==oracle", np.array_equal(g, o), "gpu==ref", np.array_equal(g, r), "maxdiff", np.abs(g - r).max(),
              "gpu ms %.3f" % ms, "ref s %.3f" % tr, "Mrays/s gpu %.1f ref %.2f" % (10 * w * h / ms / 1e3, 10 * w * h / tr / 1e6))
        # encoder
        stream = rd.draw_screen(g)
        buf = np.zeros(abi.stream_bytes(w, h), dtype=np.uint8)
        scr = abi.Screen(o.ctypes.data_as(C.POINTER(abi.Vector)), w, h)
        orc.orc_encode_stream(C.byref(scr), C.c_void_p(buf.ctypes.data))
        fused = np.array(rd.render_ansi(sc))
        print("   stream==oracle", np.array_equal(stream, buf), "fused==oracle", np.array_equal(fused, buf), rd.last_ms())
sky = S.get_skybox("uv_checker"); rd.upload_skybox(sky)
for (w, h) in [(1920, 1080), (3840, 2160)]:
    sc = S.SceneData(w, h, sky).set_time(3.7)
    for i in range(3):
        fused = rd.render_ansi(sc)
        print(w, h, "fused ms", rd.last_ms(), "Mrays/s %.1f" % (10 * w * h / rd.last_ms()[0] / 1e3))
    rd.set_scene(sc)
    ctr, F = rd.count_rows(w, h, 0, h)
    print("counters", ctr[:15], "F=%.4g flop, per sample %.1f" % (F, F / (10 * w * h)))
