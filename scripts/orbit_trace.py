"""Development aid: where the time of OrbitPipeline.stream goes (one GPU).  Timestamps of acquire / landed / written per frame."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
from terminalraytracer_b200 import abi, pipeline, renderer as R, scene as S, sharding, lib as _lib
w, h, n = 1920, 1080, 120
sky = S.get_skybox("milky_way")
rd = R.Renderer(0, sky)
times = sharding.orbit_times(360)[:n]
# 1. the C loop with the library's own pinned buffers and a Python sink that does nothing
t0 = time.perf_counter()
rd.render_orbit(S.SceneData(w, h, sky), times, lambda f, v: False)
print("trt_render_orbit + python no-op sink: %.2f ms/frame" % ((time.perf_counter() - t0) * 1e3 / n), flush=True)
# 2. trt_render_orbit_to into ONE pinned buffer region (cudaMallocHost), no ring logic
total = abi.stream_bytes(w, h)
buf = rd.L.trt_host_alloc_pinned(3 * total)
acq = _lib.FRAME_ACQUIRE(lambda f, nb, u: buf + (f % 3) * total)
snk = _lib.FRAME_SINK(lambda p, nb, f, u: 0)
arr = (C.c_double * n)(*times)
sc = S.SceneData(w, h, sky)
t0 = time.perf_counter()
rd.L.trt_render_orbit_to(C.byref(sc.c), w, h, arr, n, 0, 1, C.cast(acq, C.c_void_p), C.cast(snk, C.c_void_p), None)
print("trt_render_orbit_to, cudaMallocHost destinations, python callbacks: %.2f ms/frame" % ((time.perf_counter() - t0) * 1e3 / n), flush=True)
# 3. the same into a registered shared-memory ring without a consumer thread
ring = pipeline.OrderedFrameRing(total, n, 3, 0, 1, None, rd)
stamps = []
def acq3(f, nb, u):
    stamps.append(("acq", f, time.perf_counter()))
    ring.header[0] = f            # pretend everything has been consumed
    return ring.slot_address(f)
def snk3(p, nb, f, u):
    stamps.append(("land", f, time.perf_counter()))
    return 0
a3, s3 = _lib.FRAME_ACQUIRE(acq3), _lib.FRAME_SINK(snk3)
t0 = time.perf_counter()
rd.L.trt_render_orbit_to(C.byref(sc.c), w, h, arr, n, 0, 1, C.cast(a3, C.c_void_p), C.cast(s3, C.c_void_p), None)
print("trt_render_orbit_to, registered shm ring, no consumer: %.2f ms/frame" % ((time.perf_counter() - t0) * 1e3 / n), flush=True)
ring.close()
# 4. the full pipeline
orbit = pipeline.OrbitPipeline(rd, w, h)
orbit.stream(S.SceneData(w, h, sky), times[:8], None)
devnull = os.open(os.devnull, os.O_WRONLY)
t0 = time.perf_counter()
orbit.stream(S.SceneData(w, h, sky), times, lambda k, v: os.write(devnull, v) != len(v))
print("OrbitPipeline.stream to /dev/null: %.2f ms/frame" % ((time.perf_counter() - t0) * 1e3 / n), flush=True)
t0 = time.perf_counter()
orbit.stream(S.SceneData(w, h, sky), times, None)
print("OrbitPipeline.stream, discard: %.2f ms/frame" % ((time.perf_counter() - t0) * 1e3 / n), flush=True)
orbit.close()
rd.close()
