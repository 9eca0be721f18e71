"""Measures the five BASELINE.json configs on one GPU (kernel-only and end-to-end) and checks sampled rows against
the oracle.  Development/report aid: writes profiles/r01_configs.json.  usage: run_configs.py [out.json]"""
import ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from terminalraytracer_b200 import abi, renderer as R, scene as S, sharding
from tests import _util as U

out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/configs.json"
orc = U.load_oracle()
rd = R.Renderer(0)
results = {}


def check_rows(sc, stream, rows):
    w, h = sc.width, sc.height
    body = stream[6:-3].reshape(h, abi.row_bytes(w))
    bad = 0
    for r in rows:
        want = U.oracle_stream(orc, U.oracle_rows(orc, sc, r, r + 1))[6:-3]
        bad += int((body[r] != want).sum())
    return bad


def run(name, sc, sky, frames=3, rows=(0, 7)):
    rd.upload_skybox(sky)
    rd.set_scene(sc)
    ctr, F = rd.count_rows(sc.width, sc.height, 0, sc.height)
    k1 = []
    e2e = 0.0
    rd.render_ansi(sc)                     # warm-up: the pinned output buffer is allocated on first use
    for _ in range(frames):
        t0 = time.perf_counter()
        view = rd.render_ansi(sc)          # the C-ABI call: H2D scene, K1, K2, D2H into pinned memory
        e2e += (time.perf_counter() - t0) / frames
        k1.append(rd.last_ms()[0])
        stream = np.array(view)            # (the copy out of the pinned buffer is the test's, not the path's)
    rays = 10.0 * sc.width * sc.height
    bad = check_rows(sc, stream, [r for r in rows if r < sc.height])
    results[name] = {"width": sc.width, "height": sc.height, "spheres": sc.c.num_spheres, "k1_ms": min(k1), "e2e_ms": e2e * 1e3,
                     "Mrays_s_kernel": rays / min(k1) / 1e3, "Mrays_s_e2e": rays / e2e / 1e6, "flops_per_ray": F / rays,
                     "model_TFLOPs": F / min(k1) / 1e9, "exact_sphere_tests_pct": 100.0 * ctr[27] / max(ctr[0], 1),
                     "cull_violations": ctr[28], "oracle_rows_checked": list(rows), "mismatching_bytes": bad}
    print(name, json.dumps(results[name]))


sky_uv, sky_mw = S.get_skybox("uv_checker"), S.get_skybox("milky_way")
run("config0_default_480x280", S.SceneData(480, 280, sky_mw).set_time(3.7), sky_mw, rows=range(0, 280, 40))
run("config1_3840x2160_uv_checker", S.SceneData(3840, 2160, sky_uv).set_time(3.7), sky_uv, rows=(0, 700, 1500, 2159))
run("config2_7680x4320_milky_way", S.SceneData(7680, 4320, sky_mw).set_time(3.7), sky_mw, rows=(0, 2100, 4319))
run("config3_stress1024_3840x2160", S.SceneData(3840, 2160, sky_uv, kind="stress", num_spheres=1024).set_time(3.7), sky_uv, frames=1, rows=())
run("config3_stress1024_480x270_oracle_checked", S.SceneData(480, 270, sky_uv, kind="stress", num_spheres=1024).set_time(3.7), sky_uv, frames=2, rows=(100, 200))
# config 4: 360-frame orbit at 1920x1080 on one GPU: per-frame calls (trt_render_ansi) and the streaming sink
# (trt_render_orbit: D2H of frame k overlaps the render of frame k+1); frame-sharding across GPUs is exercised by tests / dist.py
sc = S.SceneData(1920, 1080, sky_mw)
rd.upload_skybox(sky_mw)
times = sharding.orbit_times(360)
k1 = 0.0
t0 = time.perf_counter()
for t in times:
    sc.set_time(t)
    rd.render_ansi(sc)
    k1 += rd.last_ms()[0] + rd.last_ms()[1]
wall = time.perf_counter() - t0
devnull = open(os.devnull, "wb")
t0 = time.perf_counter()
n = rd.render_orbit(S.SceneData(1920, 1080, sky_mw), times, lambda f, v: devnull.write(v) and False)
wall_stream = time.perf_counter() - t0
assert n == 360
results["config4_orbit_360x1920x1080_1gpu"] = {"frames": 360, "kernel_fps": 360 / (k1 * 1e-3), "e2e_fps": 360 / wall,
                                               "e2e_fps_streaming_sink_to_devnull": 360 / wall_stream,
                                               "Mrays_s_e2e": 360 * 10.0 * 1920 * 1080 / wall / 1e6,
                                               "Mrays_s_e2e_streaming": 360 * 10.0 * 1920 * 1080 / wall_stream / 1e6}
print("config4", json.dumps(results["config4_orbit_360x1920x1080_1gpu"]))
json.dump(results, open(out_path, "w"), indent=1)
rd.close()
