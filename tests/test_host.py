"""Host-side logic of the product (no GPU): ABI surface, scene fixtures, camera recipe, PPM ingest,
sharding.  The camera/scene helpers are compared with the reference's own functions through the
golden vectors (and live when the reference build is present)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from terminalraytracer_b200 import abi, lib as trtlib, scene as S, sharding
from tests import _util as U


def header_symbols():
    text = open(os.path.join(U.ROOT, "include", "trt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(trt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(trt):
    declared = header_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(trt, name), f"{name} declared in include/trt_b200.h but not exported"
    # and the Python binding types exactly the declared set
    assert sorted(trtlib.SIGNATURES) == declared


def test_library_is_not_initialised_without_a_gpu_call(trt):
    assert trt.trt_is_initialized() in (0, 1)


def test_camera_recipe_matches_reference_golden(trt):
    units = np.load(os.path.join(U.GOLDEN, "units.npz"))
    for t, want in zip(units["camera_t"], units["camera_frame"]):
        cam = abi.Camera()
        trt.trt_init_camera(C.byref(cam), 480, 280)
        trt.trt_orbit_camera(C.byref(cam), float(t))
        got = np.frombuffer(bytes(cam.frame), dtype=np.float64)
        assert np.array_equal(got, want), t
    dx = (C.c_double * 10)()
    dy = (C.c_double * 10)()
    trt.trt_subpixel_offsets(dx, dy)
    assert list(dx) == list(units["sub_dx"]) and list(dy) == list(units["sub_dy"])


def test_camera_recipe_matches_reference_live(trt, ref):
    for t in np.linspace(0, 40, 37):
        a, b = abi.Camera(), abi.Camera()
        trt.trt_init_camera(C.byref(a), 480, 280)
        trt.trt_init_camera(C.byref(b), 480, 280)
        trt.trt_orbit_camera(C.byref(a), float(t))
        ref.ref_orbit_camera(C.byref(b), float(t))
        assert bytes(a) == bytes(b)
    cam = abi.Camera()
    ref.init_camera(C.byref(cam))
    mine = abi.Camera()
    trt.trt_init_camera(C.byref(mine), 480, 280)
    assert bytes(cam) == bytes(mine)


def test_demo_scene_literals(trt):
    """TRT.c:1256-1306"""
    sc = S.SceneData(480, 280, S.synthetic_cubemap("colors", 8))
    c = sc.c
    assert c.num_spheres == 6 and c.num_directional_lights == 1 and c.num_point_lights == 1
    centres = [c.spheres[i].center.tup() for i in range(6)]
    assert centres == [(1, 0, 0), (0, 1, 0), (0, 0, 1), (-1, 0, 0), (0, -1, 0), (0, 0, -1)]
    assert [c.spheres[i].radius for i in range(6)] == [0.5] * 6
    assert [c.spheres[i].material.reflectivity for i in range(6)] == [1.0, 0.8, 0.8, 0.8, 0.8, 0.8]
    assert [c.spheres[i].material.color.tup() for i in range(6)] == [(1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 1, 1), (1, 0, 1), (1, 1, 0)]
    assert c.ground.point.tup() == (0, -2, 0) and c.ground.normal.tup() == (0, 1, 0)
    assert c.ground.even_material.color.tup() == (1, 1, 1) and c.ground.odd_material.color.tup() == (1, 0, 0)
    assert c.ground.even_material.reflectivity == 0.2 == c.ground.odd_material.reflectivity
    assert c.directional_lights[0].direction.tup() == (-1, -1, -1)
    assert c.point_lights[0].position.tup() == (0, 0, 0) and c.point_lights[0].intensity == 10.0
    assert c.camera.screen_distance == 1.0 and c.camera.screen_height == 5.0
    assert c.camera.screen_width == 5 * 480.0 / 280.0


def test_stress_scene_is_deterministic_and_clear_of_the_orbit(trt):
    a = S.SceneData(64, 36, S.synthetic_cubemap("colors", 8), kind="stress", num_spheres=1024)
    b = S.SceneData(64, 36, S.synthetic_cubemap("colors", 8), kind="stress", num_spheres=1024)
    assert bytes(a.spheres) == bytes(b.spheres)
    refl = set()
    for i in range(1024):
        s = a.spheres[i]
        d = np.linalg.norm(s.center.tup())
        assert not (1.99 - s.radius - 0.05 < d < 1.99 + s.radius + 0.05)
        assert 0.1 <= s.radius <= 0.5
        refl.add(s.material.reflectivity)
    assert refl == {0.0, 0.2, 0.8, 1.0}


def test_ppm_reader_matches_c_loader_and_grammar(trt, tmp_path):
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (5, 5, 3), dtype=np.uint8)
    d = tmp_path / "sky"
    d.mkdir()
    for i, name in enumerate(S.FACE_FILES):
        with open(d / name, "wb") as f:
            # GIMP-style header with a comment line, as in the reference's assets
            f.write(b"P6\n# Created by GIMP version 2.10.24 PNM plug-in\n5 5\n255\n")
            f.write(((img.astype(int) + i) % 256).astype(np.uint8).tobytes())
    sky_py = S.load_skybox_dir(str(d))
    sky_c = abi.Skybox()
    trt.trt_load_skybox_dir(C.byref(sky_c), str(d).encode())
    assert sky_c.dim == 5 == sky_py.dim
    for f in range(6):
        got = np.ctypeslib.as_array(C.cast(sky_c.colors[f], C.POINTER(C.c_ubyte)), shape=((25 + 6) * 3,))
        assert np.array_equal(got[:75].reshape(5, 5, 3), sky_py.face(f))
        assert not got[75:].any()  # dim+1 black pad texels
    trt.trt_free_skybox(C.byref(sky_c))
    assert sky_c.dim == -1


def test_reference_assets_load_identically(trt):
    d = os.path.join(U.REFERENCE_DIR, "skybox", "colors")
    if not os.path.isdir(d):
        pytest.skip("reference assets not present")
    real = S.load_skybox_dir(d)
    synth = S.synthetic_cubemap("colors", 256)
    for f in range(6):
        assert np.array_equal(real.face(f), synth.face(f))  # the synthetic stand-in IS the reference card


def test_row_bands_cover_every_row_once():
    for h in (1, 7, 280, 2160, 4320):
        for n in (1, 2, 3, 4, 8):
            bands = sharding.row_bands(h, n)
            assert bands[0][0] == 0 and bands[-1][1] == h and len(bands) == n
            for (a0, a1), (b0, b1) in zip(bands, bands[1:]):
                assert a1 == b0 and a0 <= a1
    assert sharding.row_bands(4320, 8) == [(i * 540, (i + 1) * 540) for i in range(8)]
    assert sharding.row_bands(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]


def test_weighted_row_bands_balance_cost():
    w = [1.0] * 100 + [5.0] * 100
    bands = sharding.row_bands(200, 4, w)
    assert bands[0][0] == 0 and bands[-1][1] == 200
    costs = [sum(w[a:b]) for a, b in bands]
    assert max(costs) - min(costs) <= 10.0
    for (a0, a1), (b0, b1) in zip(bands, bands[1:]):
        assert a1 == b0


def test_band_byte_ranges_tile_the_stream():
    w, h = 37, 23
    bands = sharding.row_bands(h, 5)
    pos = abi.HOME_BYTES
    for b in bands:
        b0, b1 = sharding.band_byte_range(w, b)
        assert b0 == pos
        pos = b1
    assert pos + abi.TAIL_NULS == abi.stream_bytes(w, h)


def test_frame_sharding():
    assert sharding.frames_for_rank(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((sharding.frames_for_rank(360, r, 8) for r in range(8)), [])) == list(range(360))
    ts = sharding.orbit_times(360)
    assert ts[0] == 0.0 and abs(ts[-1] - (20.0 - 20.0 / 360)) < 1e-12


def test_synthetic_skyboxes_are_deterministic():
    a = S.synthetic_cubemap("milky_way", 64)
    b = S.synthetic_cubemap("milky_way", 64)
    for f in range(6):
        assert np.array_equal(a.face(f), b.face(f))
    stars = sum(int((a.face(f).max(axis=2) > 50).sum()) for f in range(6))
    assert 0.001 < stars / (6 * 64 * 64) < 0.01


def test_sub_bands_partition_a_band_exactly():
    """pieces of a rank's band (pushed to rank 0 one by one): contiguous, non-empty, covering, cost-proportional"""
    from terminalraytracer_b200 import sharding
    for band in [(0, 1), (0, 2), (7, 19), (100, 1100), (5, 5)]:
        for pieces in (1, 2, 3, 7, (0.7, 0.3), (0.5, 0.3, 0.2)):
            sb = sharding.sub_bands(band, pieces)
            if band[1] == band[0]:
                assert sb == []
                continue
            assert sb[0][0] == band[0] and sb[-1][1] == band[1]
            assert all(a[1] == b[0] for a, b in zip(sb, sb[1:])) and all(b > a for a, b in sb)
    assert sharding.sub_bands((0, 100), (0.7, 0.3)) == [(0, 70), (70, 100)]
    w = [1.0] * 50 + [9.0] * 50                     # the second half of the rows costs 9x more
    (a0, a1), (b0, b1) = sharding.sub_bands((0, 100), (0.5, 0.5), w)
    assert a0 == 0 and a1 == b0 and b1 == 100 and 70 <= a1 <= 80


def test_reweight_feedback_balances_bands_within_a_few_frames():
    """adaptive bands (FramePipeline adapt=True): a poor first estimate, then the ranks' measured times.  The true cost
    profile is never seen row by row — only one total per band and frame — and the bands still converge."""
    import numpy as np
    rng = np.random.default_rng(7)
    h, world = 4320, 8
    rows = np.arange(h)
    true = 1.0 + 4.0 * np.exp(-((rows - 2600) / 500.0) ** 2) + 2.0 * (rows > 3000)      # cheap sky on top, spheres, ground
    weights = np.ones(h)                                                                 # knows nothing
    imbalance = []
    for frame in range(6):
        bands = sharding.row_bands(h, world, weights)
        assert bands[0][0] == 0 and bands[-1][1] == h and all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
        times = [float(true[r0:r1].sum()) * (1.0 + 0.01 * rng.standard_normal()) for r0, r1 in bands]   # 1 % timing noise
        imbalance.append(max(times) / (sum(times) / world))
        weights = sharding.reweight(weights, bands, times)
    assert imbalance[0] > 1.5
    assert imbalance[-1] < 1.05, imbalance


def test_reweight_keeps_unmeasured_bands_and_handles_empty_ones():
    import numpy as np
    w = [1.0] * 10
    out = sharding.reweight(w, [(0, 5), (5, 5), (5, 10)], [2.0, 0.0, 6.0])
    assert np.allclose(out[:5].sum(), 2.0) and np.allclose(out[5:].sum(), 6.0)
    out = sharding.reweight(w, [(0, 5), (5, 10)], [0.0, 0.0])          # nothing measured: unchanged
    assert np.allclose(out, w)
    out = sharding.reweight(w, [(0, 4), (4, 10)], [8.0, 0.0])          # rank 1 has no measurement: scaled like the rest
    assert np.allclose(out[:4], 2.0) and np.allclose(out[4:], 2.0)


def test_row_bands_numpy_cuts_match_the_sequential_definition():
    import numpy as np
    rng = np.random.default_rng(3)
    for _ in range(20):
        h, world = int(rng.integers(1, 300)), int(rng.integers(1, 9))
        w = rng.random(h) * (rng.random(h) > 0.2)
        want, r, acc, total = [], 0, 0.0, float(np.cumsum(w)[-1])
        if total <= 0:
            continue
        for i in range(world):
            r1 = r
            while r1 < h and (acc + w[r1] <= total * (i + 1) / world or i == world - 1):
                acc += w[r1]
                r1 += 1
            want.append((r, r1))
            r = r1
        assert sharding.row_bands(h, world, w) == want


def test_shared_host_stream_availability_probe():
    """bench.py falls back to the rank-0 copy when /dev/shm cannot hold the stream"""
    from terminalraytracer_b200 import pipeline
    assert pipeline.SharedHostStream.available(1 << 16) in (True, False)
    assert pipeline.SharedHostStream.available(1 << 62) is False


def test_committed_bench_line_has_the_contract_keys():
    """the JSON line bench.py printed on the B200 (profiles/r02_bench_n1.json) carries every key of the bench contract"""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "profiles", "r02_bench_n1.json")) as f:
        line = json.loads(f.read())
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["metric"] == "Mrays/s" and line["unit"] == "Mrays/s" and line["higher_is_better"] is True
    assert line["n_gpus"] == 1 and line["warmup"] >= 3 and line["gpu_launches"] == 4 * line["steps"]      # k_tile_certs, K1, K2, k_stream_frame
    assert "workload" in line["config"] and "model" not in line["config"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in line["roofline"], key
    assert abs(line["roofline"]["frac"] - line["roofline"]["achieved"] / line["roofline"]["peak"]) < 1e-9
    for key in ("value", "unit", "cores", "kind", "sample"):
        assert key in line["cpu_baseline"], key
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert key in line["e2e"], key
    assert line["e2e"]["d2h_bytes_per_step"] == abi.stream_bytes(7680, 4320)
    assert line["e2e"]["value"] < line["value"]                       # host copies inside the timed region
    assert line["clocks"]["reasons"] == [] or "sw_power_cap" in line["clocks"]["reasons"]
    # value = primary rays per second over the timed steps
    assert abs(line["value"] - 10.0 * 7680 * 4320 / (line["ms_per_step"] * 1e-3) / 1e6) < 1e-6 * line["value"]


def test_bench_stress_scene_restatement_matches_the_library(trt):
    """bench.py's reference arm builds the 1024-sphere scene without loading libtrt_b200.so; the restatement must be the same bytes"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(U.ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    sc = S.SceneData(64, 36, S.synthetic_cubemap("colors", 16), kind="stress", num_spheres=1024)
    assert bytes(bench._stress_spheres(1024)) == bytes(sc.spheres)


def test_reference_arm_never_maps_the_product_library():
    """bench.py --impl reference: the scene comes from the reference's own init_camera / literals / camera recipe and the timed
    call is its project_scene; libtrt_b200.so must not be mapped into that process (reference-arm hygiene)"""
    if not U.have_reference_build():
        pytest.skip("oracle/_ref not built")
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import bench\n"
            "cfg = dict(bench.CONFIGS['demo8k']); cfg['cpu_rows'] = [2000]\n"
            "cpu = bench.ReferenceCPU(cfg); s = cpu.one_pass(); assert cpu.kind == 'reference' and s > 0 and cpu.px.any()\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'libtrt_ref_rows.so' in maps and 'libtrt_b200' not in maps, maps\n"
            "print('clean')") % U.ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "clean" in r.stdout, r.stderr[-2000:]


def test_pose_camera_generalises_the_orbit_recipe(trt):
    """trt_pose_camera (keyboard-driven cameras) with the reference's angles is the reference's orbit camera, bit for bit"""
    import math
    for t in (0.0, 0.5, 3.7, 19.99, 123.456):
        a, b = abi.Camera(), abi.Camera()
        trt.trt_init_camera(C.byref(a), 480, 280)
        trt.trt_init_camera(C.byref(b), 480, 280)
        trt.trt_orbit_camera(C.byref(a), t)
        trt.trt_pose_camera(C.byref(b), 2.0 * math.pi * t * -0.03, 2.0 * math.pi * t * 0.05, 1.99)
        assert bytes(a) == bytes(b), t
