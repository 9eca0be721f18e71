"""compute-sanitizer workload: small frames through EVERY kernel of libtrt_b200 —
  k_render<COUNT 0|1, CULL 0|1|2, LIGHTS 0|1> (demo scene, 200-sphere k-d scene, certificates off, 2 + 3 lights; the counting
  flavour of each), k_tile_certs, the fused encode epilogue at four byte alignments of the stream, k_encode<double> and
  k_encode<uchar4>, k_stream_frame, the orbit sink with its double-buffered copies, the probes, k_signal / k_wait_flags.
Run under each tool and keep the ERROR SUMMARY lines (scripts/run_sanitizer.sh -> profiles/r02_sanitizer.txt):
    compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck python scripts/sanitize_small.py
No torch: ctypes + numpy only, so the tool instruments nothing but our kernels."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from terminalraytracer_b200 import abi, renderer as R, scene as S

sky = S.synthetic_cubemap("uv_gradient", 64)
rd = R.Renderer(0, sky)
L = rd.L


def many_lights(w, h):
    sc = S.SceneData(w, h, sky).set_time(3.7)
    rng = np.random.default_rng(5)
    sc.dls = (abi.DirectionalLight * 2)()
    sc.pls = (abi.PointLight * 3)()
    for i in range(2):
        sc.dls[i] = abi.DirectionalLight(abi.Vector(*rng.normal(size=3)), abi.Vector(*rng.uniform(0.2, 1, 3)))
    for i in range(3):
        sc.pls[i] = abi.PointLight(abi.Vector(*rng.uniform(-3, 3, 3)), abi.Vector(*rng.uniform(0.2, 1, 3)), float(rng.uniform(1, 20)))
    sc.c.directional_lights = C.cast(sc.dls, C.POINTER(abi.DirectionalLight))
    sc.c.num_directional_lights = 2
    sc.c.point_lights = C.cast(sc.pls, C.POINTER(abi.PointLight))
    sc.c.num_point_lights = 3
    return sc


scenes = {
    "demo <.,1,1>": S.SceneData(45, 27, sky).set_time(3.7),
    "demo ragged <.,1,1>": S.SceneData(9, 5, sky).set_time(0.0),
    "200 spheres <.,2,1>": S.SceneData(34, 20, sky, kind="stress", num_spheres=200).set_time(3.7),
    "2+3 lights <.,1,0>": many_lights(40, 22),
}
for name, sc in scenes.items():
    a = rd.project_scene(sc)                      # k_render<0,CULL,LIGHTS> (+ k_tile_certs), FP64 framebuffer
    b = np.array(rd.render_ansi(sc))              # quantised cells, k_encode<uchar4>, k_stream_frame
    d = rd.draw_screen(a)                         # k_encode<double>
    rd.set_scene(sc)
    ctr, _ = rd.count_rows(sc.width, sc.height, 0, sc.height)      # k_render<1,CULL,LIGHTS>
    L.trt_set_cull(0)
    c = rd.project_scene(sc)                      # k_render<0,0,LIGHTS>
    rd.set_scene(sc)
    ctr0, _ = rd.count_rows(sc.width, sc.height, 0, sc.height)     # k_render<1,0,LIGHTS>
    L.trt_set_cull(1)
    print(name, "cull off same:", np.array_equal(a, c), "violations", ctr[28], "stream == drop-in:", np.array_equal(b, d), flush=True)

# fused epilogue, four alignments, two bands each
sc = scenes["demo <.,1,1>"]
want = np.array(rd.render_ansi(sc)).copy()
total = abi.stream_bytes(sc.width, sc.height)
dev = L.trt_device_alloc(total + 64)
rd.set_scene(sc)
for shift in range(4):
    L.trt_render_rows_ansi_device(sc.width, sc.height, 0, 11, C.c_void_p(dev + shift))
    L.trt_render_rows_ansi_device(sc.width, sc.height, 11, sc.height, C.c_void_p(dev + shift))
    L.trt_stream_frame_device(C.c_void_p(dev + shift), sc.width, sc.height)
    got = np.zeros(total, dtype=np.uint8)
    L.trt_copy_to_host(got.ctypes.data, C.c_void_p(dev + shift), total)
    print("fused shift", shift, "same bytes:", np.array_equal(got, want), flush=True)

# step flags (multi-GPU completion): signal three ranks' flags, wait for them
flags = L.trt_device_alloc(512)
zero = np.zeros(128, dtype=np.uint32)
L.trt_copy_to_device(C.c_void_p(flags), zero.ctypes.data, 512)
for r in range(3):
    L.trt_signal_step(C.c_void_p(flags + 4 * r), 7, 0)
L.trt_wait_steps(C.c_void_p(flags), 3, 7, 0)
L.trt_synchronize()
back = np.zeros(128, dtype=np.uint32)
L.trt_copy_to_host(back.ctypes.data, C.c_void_p(flags), 512)
print("flags", back[:3].tolist(), "timeout", int(back[32]), flush=True)
L.trt_device_free(C.c_void_p(flags))
L.trt_device_free(C.c_void_p(dev))

# orbit sink (double-buffered device/host buffers, async scene uploads)
got = []
rd.render_orbit(S.SceneData(32, 18, sky), [0.0, 1.0, 2.0, 3.0, 4.0], lambda f, v: got.append(int(v.sum())) and False)
print("orbit", len(got), flush=True)

# probes
rays = np.random.default_rng(1).normal(size=(64, 6))
out = np.zeros((64, 11))
L.trt_probe_trace_ray(C.byref(sc.c), rays.ctypes.data, 64, out.ctypes.data)
geom = np.abs(np.random.default_rng(2).normal(size=(64, 4))) + 0.1
o4 = np.zeros((64, 4))
L.trt_probe_sphere(rays.ctypes.data, geom.ctypes.data, 64, o4.ctypes.data)
L.trt_probe_plane(C.byref(sc.c), rays.ctypes.data, 64, o4.ctypes.data)
surf = np.ascontiguousarray(out[:, 1:10])
o3 = np.zeros((64, 3))
L.trt_probe_lighting(C.byref(sc.c), surf.ctypes.data, 64, o3.ctypes.data)
dirs = np.ascontiguousarray(rays[:, 3:])
o5 = np.zeros((64, 5), dtype=np.int32)
L.trt_probe_skybox(dirs.ctypes.data, 64, o5.ctypes.data)
print("probes ok", flush=True)
# a few random scenes as well (scripts/fuzz_parity.py: sphere counts across the chunk and cluster limits, tilted grounds, many lights)
import importlib.util
spec = importlib.util.spec_from_file_location("fuzz_parity", os.path.join(os.path.dirname(os.path.abspath(__file__)), "fuzz_parity.py"))
fuzz = importlib.util.module_from_spec(spec)
spec.loader.exec_module(fuzz)
rng = np.random.default_rng(7)
for k in range(12):
    fsc = fuzz.random_scene(rng, sky)
    rd.project_scene(fsc)
    np.array(rd.render_ansi(fsc))
print("fuzz scenes ok", flush=True)
# self-checking build (-DTRT_BOUNDS_CHECK): the kernels' own index checks
counts = (C.c_uint * 32)()
compiled = L.trt_debug_bounds(counts)
print("bounds checks compiled in:", bool(compiled), " violations per site (render 0-15, encode 16-31):", list(counts), flush=True)
rd.close()
