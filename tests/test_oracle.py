"""Pins the oracle (oracle/trt_oracle.c, our C restatement of the reference's render path):
  * against tests/golden/ (outputs of the reference build on seeded inputs, make_golden.py), and
  * live against oracle/_ref/libtrt_ref.so when it has been built (needs /root/reference).
Everything is compared bit for bit.  CPU only."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from terminalraytracer_b200 import abi, scene as S
from tests import _util as U
from tests.golden import make_golden as G


@pytest.fixture(scope="module")
def units():
    return np.load(os.path.join(U.GOLDEN, "units.npz"))


@pytest.fixture(scope="module")
def frames():
    return np.load(os.path.join(U.GOLDEN, "frames.npz"))


@pytest.mark.parametrize("case", G.frame_cases(), ids=lambda c: c[0])
def test_oracle_frames_match_reference_golden(orc, frames, case):
    sc = G.make_scene(case)
    got = U.cpu_render(orc, "orc_project_scene", sc)
    assert np.array_equal(got, frames[case[0]])


def test_oracle_sphere_known_answers(orc, units):
    rays, geom, want = units["sphere_rays"], units["sphere_geom"], units["sphere_out"]
    assert 0.05 < want[:, 0].mean() < 0.95  # both outcomes are exercised
    for i in range(len(rays)):
        r = abi.Ray(abi.Vector(*rays[i, :3]), abi.Vector(*rays[i, 3:]))
        s = abi.Sphere(abi.Vector(*geom[i, :3]), geom[i, 3], abi.Material())
        p = abi.Vector(0, 0, 0)
        hit = orc.orc_hit_sphere(C.byref(r), C.byref(s), C.byref(p), None)
        assert hit == int(want[i, 0])
        if hit:
            assert p.tup() == tuple(want[i, 1:])


def test_oracle_plane_known_answers(orc, units):
    rays, want = units["plane_rays"], units["plane_out"]
    plane = abi.Plane(abi.Vector(0, -2, 0), abi.Vector(0, 1, 0), abi.Material(), abi.Material())
    for i in range(len(rays)):
        r = abi.Ray(abi.Vector(*rays[i, :3]), abi.Vector(*rays[i, 3:]))
        p = abi.Vector(0, 0, 0)
        hit = orc.orc_hit_plane(C.byref(r), C.byref(plane), C.byref(p), None)
        assert hit == int(want[i, 0])
        if hit:
            assert p.tup() == tuple(want[i, 1:])


def index_coded_skybox(dim=64):
    planes = []
    for f in range(6):
        idx = np.arange(dim * dim)
        img = np.stack([idx & 255, (idx >> 8) & 255, np.full_like(idx, f * 40) + (idx >> 16)], axis=1).astype(np.uint8)
        planes.append(img.reshape(dim, dim, 3))
    return S.SkyboxData(planes)


def test_oracle_skybox_known_answers(orc, units):
    sky = index_coded_skybox()
    dirs, want = units["sky_dirs"], units["sky_rgb"]
    faces = set()
    for i in range(len(dirs)):
        v = abi.Vector(*dirs[i])
        c = abi.Color()
        face = C.c_int()
        idx = C.c_long()
        orc.orc_sky_texel(C.byref(sky.c), C.byref(v), C.byref(c), C.byref(face), C.byref(idx))
        assert (c.r, c.g, c.b) == tuple(want[i]), (i, dirs[i])
        faces.add(face.value)
    assert faces == set(range(6))
    # the pad texels (index >= dim*dim) are black and were reached by the axis-aligned directions
    assert (want == 0).all(axis=1).any()


def test_oracle_trace_and_lighting_known_answers(orc, units):
    sky = S.synthetic_cubemap("uv_gradient", 64)
    sc = S.SceneData(64, 36, sky).set_time(3.7)
    rays, want, lit = units["trace_rays"], units["trace_out"], units["lighting_out"]
    kinds = set()
    for i in range(len(rays)):
        r = abi.Ray(abi.Vector(*rays[i, :3]), abi.Vector(*rays[i, 3:]))
        p, n, m = abi.Vector(), abi.Vector(), abi.Material()
        kind = orc.orc_closest_hit(C.byref(sc.c), C.byref(r), C.byref(p), C.byref(n), C.byref(m), None)
        got = (kind, p.x, p.y, p.z, n.x, n.y, n.z, m.color.x, m.color.y, m.color.z, m.reflectivity)
        assert got == tuple(want[i]), i
        kinds.add(kind)
        if kind:
            orc.orc_light_surface(C.byref(sc.c), C.byref(p), C.byref(n), C.byref(m), None)
            assert m.color.tup() == tuple(lit[i]), i
    assert kinds == {0, 1, 2}


def test_oracle_subpixel_offsets(orc, units):
    dx = (C.c_double * 10)()
    dy = (C.c_double * 10)()
    orc.orc_subpixel_offsets(dx, dy)
    assert list(dx) == list(units["sub_dx"]) and list(dy) == list(units["sub_dy"])
    # the pattern SURVEY §7.3 H3 describes
    assert np.allclose(list(dx), [0, .1, .2, .3, .4, .5, .4, .3, .2, .1])
    assert np.allclose(list(dy), [0.05 * k for k in range(10)])


def test_oracle_stream_matches_reference_golden(orc):
    with open(os.path.join(U.GOLDEN, "streams.json")) as f:
        gold = json.load(f)
    for key, g in gold.items():
        if "skybox" not in g:
            continue
        w, h = (int(v) for v in key.split("_")[0].split("x"))
        sky = S.synthetic_cubemap(g["skybox"], g["dim"])
        sc = S.SceneData(w, h, sky).set_time(g["t"])
        px = U.cpu_render(orc, "orc_project_scene", sc)
        assert U.sha(px) == g["pixels_sha256"], key
        stream = U.oracle_stream(orc, px)
        assert stream.size == g["stream_bytes"] == abi.stream_bytes(w, h)
        assert U.sha(stream) == g["stream_sha256"], key
    px = U.random_encoder_pixels()
    g = gold["480x280_random_pixels"]
    assert U.sha(px) == g["pixels_sha256"]
    assert U.sha(U.oracle_stream(orc, px)) == g["stream_sha256"]


def test_oracle_stream_layout(orc):
    """home + H x (W cells + newline) + 3 NULs (TRT.c:1102-1104, 1130, 1171)"""
    px = np.zeros((2, 3, 3))
    px[0, 0] = (1.0, 0.5, 0.999)
    s = U.oracle_stream(orc, px).tobytes()
    assert s[:6] == b"\033[0;0H"
    assert s[6:31] == b"\033[48;2;255;127;254m  \033[0m"
    assert s[6 + 75] == ord("\n")
    assert s[-3:] == b"\0\0\0" and len(s) == 9 + (25 * 3 + 1) * 2


def test_oracle_digits(orc, units):
    """byte_to_digits incl. out-of-range ints, via one-pixel streams"""
    vals, want = units["digit_values"], units["digit_chars"]
    for v, d in zip(vals, want):
        if not (0 <= v <= 2000):
            continue  # negative ints cannot be produced from a double without going through (int)
        px = np.full((1, 1, 3), (v + 0.5) / 255.0)
        s = U.oracle_stream(orc, px)
        if int(px[0, 0, 0] * 255) != v:
            continue
        assert bytes(s[6 + 7:6 + 10]) == bytes(d), v


def test_oracle_row_ranges_compose(orc):
    sky = S.synthetic_cubemap("uv_gradient", 64)
    sc = S.SceneData(40, 22, sky).set_time(5.0)
    full = U.cpu_render(orc, "orc_project_scene", sc)
    parts = [U.oracle_rows(orc, sc, a, b) for a, b in ((0, 5), (5, 6), (6, 22))]
    assert np.array_equal(np.concatenate(parts), full)


def test_oracle_counters_and_flop_model(orc):
    sky = S.synthetic_cubemap("colors", 256)
    sc = S.SceneData(48, 28, sky).set_time(0.0)
    ctr = U.Counters()
    U.oracle_rows(orc, sc, 0, 28, ctr)
    assert ctr.pixels == 48 * 28 and ctr.samples == 10 * 48 * 28
    assert ctr.trace_calls == ctr.bounce_iters + 2 * ctr.lighting_calls  # one shadow query per light (1 + 1)
    assert ctr.sphere_tests == 6 * ctr.trace_calls and ctr.plane_tests == ctr.trace_calls
    assert sum(ctr.bounce_hist) == ctr.samples
    assert ctr.sky_lookups == ctr.trace_calls - ctr.trace_hits
    f = orc.orc_model_flops(C.byref(ctr)) / ctr.samples
    assert 800 < f < 1600  # SURVEY §8d measured 1016-1337 flop/sample on this scene family


# ---- live comparison with the compiled reference ------------------------------------------------

LIVE_CASES = [("colors", 256, 120, 70, 0.0), ("uv_gradient", 64, 120, 70, 3.7), ("milky_way", 128, 96, 54, 17.2),
              ("uv_gradient", 64, 1, 1, 1.0), ("uv_gradient", 64, 5, 3, 2.0)]


@pytest.mark.parametrize("case", LIVE_CASES, ids=lambda c: f"{c[0]}_{c[2]}x{c[3]}_t{c[4]}")
def test_oracle_equals_reference_live(orc, ref, case):
    skyname, dim, w, h, t = case
    sc = S.SceneData(w, h, S.synthetic_cubemap(skyname, dim)).set_time(t)
    assert np.array_equal(U.cpu_render(orc, "orc_project_scene", sc), U.cpu_render(ref, "project_scene", sc))


def test_oracle_equals_reference_live_real_assets(orc, ref):
    d = os.path.join(U.REFERENCE_DIR, "skybox", "uv_checker")
    if not os.path.isdir(d):
        pytest.skip("reference skybox assets not present")
    sc = S.SceneData(96, 56, S.load_skybox_dir(d)).set_time(3.7)
    assert np.array_equal(U.cpu_render(orc, "orc_project_scene", sc), U.cpu_render(ref, "project_scene", sc))


def test_oracle_stress_scene_equals_reference_live(orc, ref):
    sc = S.SceneData(40, 24, S.synthetic_cubemap("uv_gradient", 64), kind="stress", num_spheres=200).set_time(3.7)
    assert np.array_equal(U.cpu_render(orc, "orc_project_scene", sc), U.cpu_render(ref, "project_scene", sc))


def test_reference_row_range_build_is_equivalent(ref):
    """the sed-patched build used to spread golden generation over cores renders the same pixels"""
    if not os.path.exists(U.REF_ROWS_SO):
        pytest.skip("row-range reference build missing")
    rows = U.load_reference(U.REF_ROWS_SO)
    rows.ref_project_rows.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.Screen), C.c_int, C.c_int]
    sc = S.SceneData(60, 34, S.synthetic_cubemap("uv_gradient", 64)).set_time(3.7)
    full = U.cpu_render(ref, "project_scene", sc)
    px = np.zeros_like(full)
    scr = U.screen_for(px)
    for a, b in ((0, 10), (10, 34)):
        rows.ref_project_rows(C.byref(sc.c), C.byref(scr), a, b)
    assert np.array_equal(px, full)


def test_abi_layout_matches_reference(ref):
    assert ref.ref_sizeof_scene() == C.sizeof(abi.Scene)
    assert ref.ref_sizeof_sphere() == C.sizeof(abi.Sphere)
    assert ref.ref_sizeof_plane() == C.sizeof(abi.Plane)
    assert ref.ref_sizeof_camera() == C.sizeof(abi.Camera)
    assert ref.ref_sizeof_skybox() == C.sizeof(abi.Skybox)
    assert ref.ref_sizeof_screen() == C.sizeof(abi.Screen)
    assert ref.ref_offsetof_scene_camera() == abi.Scene.camera.offset
    assert ref.ref_offsetof_scene_skybox() == abi.Scene.skybox.offset
    assert ref.ref_offsetof_scene_ground() == abi.Scene.ground.offset
    assert ref.ref_offsetof_scene_point_lights() == abi.Scene.point_lights.offset
    assert ref.ref_screenbuffer_bytes() == abi.stream_bytes(480, 280) == 3360289
