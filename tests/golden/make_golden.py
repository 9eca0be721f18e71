"""Writes the golden fixtures under tests/golden/ by RUNNING THE REFERENCE (oracle/_ref/libtrt_ref.so,
the unmodified /root/reference/TerminalRayTracer.c compiled by oracle/Makefile).  Run in the build
container, where /root/reference exists:      python tests/golden/make_golden.py

The reference ships no tests or vectors of its own; these are its outputs on seeded
inputs, so that the oracle restatement and the CUDA path can be pinned on boxes without /root/reference.

  frames.npz  raw FP64 framebuffers of project_scene (TRT.c:966) for small frames
  units.npz   known answers of ray_intersects_sphere/plane, get_skybox_color, trace_ray, apply_lighting,
              byte_to_digits, triangle-wave offsets and the orbit camera
  streams.json sha256 of buffered_draw_screen's byte stream (TRT.c:1142) at the reference's own 480x280
              and at 96x56 (sed-resized build), plus sha256 of larger framebuffers
"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from terminalraytracer_b200 import abi, scene as S  # noqa: E402
from tests import _util as U  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def frame_cases():
    # (name, skybox name, skybox dim or None, width, height, t, kind, nspheres)
    return [
        ("demo_colors_t0", "colors", 256, 64, 36, 0.0, "demo", 0),
        ("demo_colors_t3p7", "colors", 256, 64, 36, 3.7, "demo", 0),
        ("demo_uvgrad_t0", "uv_gradient", 64, 64, 36, 0.0, "demo", 0),
        ("demo_uvgrad_t3p7", "uv_gradient", 64, 64, 36, 3.7, "demo", 0),
        ("demo_milky_t11", "milky_way", 128, 64, 36, 11.0, "demo", 0),
        ("demo_ragged_t7", "uv_gradient", 64, 37, 23, 7.3, "demo", 0),
        ("stress40_t3p7", "uv_gradient", 64, 48, 27, 3.7, "stress", 40),
    ]


def make_scene(case):
    name, skyname, dim, w, h, t, kind, nsph = case
    sky = S.synthetic_cubemap(skyname, dim)
    sc = S.SceneData(w, h, sky, kind=kind, num_spheres=nsph if nsph else 1024)
    sc.set_time(t)
    return sc


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    ref = U.load_reference()
    rng = np.random.default_rng(20261018)

    # ---- frames ---------------------------------------------------------------------------------
    frames = {}
    for case in frame_cases():
        sc = make_scene(case)
        frames[case[0]] = U.cpu_render(ref, "project_scene", sc)
        print(case[0], frames[case[0]].mean())
    np.savez_compressed(os.path.join(OUT, "frames.npz"), **frames)

    # ---- unit vectors ---------------------------------------------------------------------------
    units = {}
    # camera poses
    ts = np.array([0.0, 0.5, 3.7, 10.0, 19.99, 123.456])
    cams = np.zeros((len(ts), 12))
    for i, t in enumerate(ts):
        cam = abi.Camera()
        ref.ref_orbit_camera(C.byref(cam), float(t))
        cams[i] = np.frombuffer(bytes(cam.frame), dtype=np.float64)
    units["camera_t"] = ts
    units["camera_frame"] = cams
    dx = (C.c_double * 10)()
    dy = (C.c_double * 10)()
    ref.ref_subpixel_offsets(dx, dy)
    units["sub_dx"] = np.array(dx[:])
    units["sub_dy"] = np.array(dy[:])

    # ray_intersects_sphere: random + tangent/inside/behind cases
    n = 4000
    rays = np.zeros((n, 6))
    rays[:, :3] = rng.uniform(-3, 3, (n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[::7] *= rng.uniform(0.1, 4.0, (len(d[::7]), 1))  # un-normalised directions: a = d.d is not assumed 1
    rays[:, 3:] = d
    sph = np.zeros((n, 4))
    sph[:, :3] = rng.uniform(-2, 2, (n, 3))
    sph[:, 3] = rng.uniform(0.1, 1.5, n)
    aim = sph[::2, :3] - rays[::2, :3] + rng.normal(size=(n // 2, 3)) * sph[::2, 3:4] * 0.7   # half the rays aim near the sphere
    rays[::2, 3:] = aim / np.linalg.norm(aim, axis=1, keepdims=True) * rng.uniform(0.2, 3.0, (n // 2, 1))
    rays[:200, :3] = sph[:200, :3] + 0.3 * sph[:200, 3:4] * d[:200]  # origins inside the sphere
    res = np.zeros((n, 4))
    for i in range(n):
        r = abi.Ray(abi.Vector(*rays[i, :3]), abi.Vector(*rays[i, 3:]))
        s = abi.Sphere(abi.Vector(*sph[i, :3]), sph[i, 3], abi.Material())
        p = abi.Vector(0, 0, 0)
        hit = ref.ray_intersects_sphere(C.byref(r), C.byref(s), C.byref(p))
        res[i] = (hit, p.x, p.y, p.z) if hit else (0, 0, 0, 0)
    units["sphere_rays"], units["sphere_geom"], units["sphere_out"] = rays, sph, res

    # ray_intersects_plane
    n = 2000
    rays = np.zeros((n, 6))
    rays[:, :3] = rng.uniform(-5, 5, (n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:100, 1] = rng.uniform(-2e-5, 2e-5, 100)  # around the |denom| > 1e-5 guard
    rays[:, 3:] = d
    plane = abi.Plane(abi.Vector(0, -2, 0), abi.Vector(0, 1, 0), abi.Material(), abi.Material())
    res = np.zeros((n, 4))
    for i in range(n):
        r = abi.Ray(abi.Vector(*rays[i, :3]), abi.Vector(*rays[i, 3:]))
        p = abi.Vector(0, 0, 0)
        hit = ref.ray_intersects_plane(C.byref(r), C.byref(plane), C.byref(p))
        res[i] = (hit, p.x, p.y, p.z) if hit else (0, 0, 0, 0)
    units["plane_rays"], units["plane_out"] = rays, res

    # get_skybox_color on an index-coded cubemap: r = idx & 255, g = (idx >> 8) & 255, b = face*40 + (idx >> 16)
    dim = 64
    planes = []
    for f in range(6):
        idx = np.arange(dim * dim)
        img = np.stack([idx & 255, (idx >> 8) & 255, np.full_like(idx, f * 40) + (idx >> 16)], axis=1).astype(np.uint8)
        planes.append(img.reshape(dim, dim, 3))
    sky = S.SkyboxData(planes)
    sc = S.SceneData(64, 36, sky)
    n = 6000
    dirs = rng.normal(size=(n, 3))
    dirs[:600] = np.round(dirs[:600])            # axis-aligned / diagonal directions: face ties, u or v = +-0.5
    dirs[600:900] *= 1e-3                        # short vectors (the length > 1e-4 guard of normalize_vector)
    dirs[np.all(dirs[:900] == 0, axis=1).nonzero()[0]] = (1.0, 0.0, 0.0)
    out = np.zeros((n, 3), dtype=np.uint8)
    for i in range(n):
        v = abi.Vector(*dirs[i])
        c = abi.Color()
        ref.get_skybox_color(C.byref(sc.c), C.byref(v), C.byref(c))
        out[i] = (c.r, c.g, c.b)
    units["sky_dirs"], units["sky_rgb"] = dirs, out

    # trace_ray and apply_lighting on the demo scene at t=3.7 (uv_gradient skybox, dim 64)
    sky2 = S.synthetic_cubemap("uv_gradient", 64)
    sc = S.SceneData(64, 36, sky2).set_time(3.7)
    n = 4000
    rays = np.zeros((n, 6))
    rays[:, :3] = rng.uniform(-2.5, 2.5, (n, 3))
    rays[:, 1] = np.abs(rays[:, 1])
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 3:] = d
    tr = np.zeros((n, 11))
    lit = np.zeros((n, 3))
    for i in range(n):
        r = abi.Ray(abi.Vector(*rays[i, :3]), abi.Vector(*rays[i, 3:]))
        p, nrm, m = abi.Vector(), abi.Vector(), abi.Material()
        kind = ref.trace_ray(C.byref(sc.c), C.byref(r), C.byref(p), C.byref(nrm), C.byref(m))
        tr[i] = (kind, p.x, p.y, p.z, nrm.x, nrm.y, nrm.z, m.color.x, m.color.y, m.color.z, m.reflectivity)
        if kind != 0:
            view = abi.Vector(-d[i, 0], -d[i, 1], -d[i, 2])
            ref.apply_lighting(C.byref(sc.c), C.byref(p), C.byref(view), C.byref(nrm), C.byref(m))
            lit[i] = m.color.tup()
    units["trace_rays"], units["trace_out"], units["lighting_out"] = rays, tr, lit

    # byte_to_digits for every int the encoder can meet and a few it should not
    vals = np.array(list(range(0, 256)) + [256, 300, 999, 1000, 1234, -1, -5, -17, -255], dtype=np.int32)
    digs = np.zeros((len(vals), 3), dtype=np.uint8)
    for i, v in enumerate(vals):
        buf = C.create_string_buffer(4)
        ref.byte_to_digits(int(v), buf)
        digs[i] = np.frombuffer(buf.raw[:3], dtype=np.uint8)
    units["digit_values"], units["digit_chars"] = vals, digs
    np.savez_compressed(os.path.join(OUT, "units.npz"), **units)

    # ---- byte streams and bigger frames, by hash --------------------------------------------------
    streams = {}
    for skyname, dim, t in (("colors", 256, 0.0), ("uv_gradient", 64, 3.7), ("milky_way", 128, 11.0)):
        sky = S.synthetic_cubemap(skyname, dim)
        sc = S.SceneData(480, 280, sky).set_time(t)
        px = U.cpu_render(ref, "project_scene", sc)
        buf = np.zeros(abi.stream_bytes(480, 280) + 64, dtype=np.uint8)
        nbytes = ref.ref_draw_screen_bytes(C.byref(U.screen_for(px)), U.VP(buf.ctypes.data), buf.size)
        assert nbytes == abi.stream_bytes(480, 280), nbytes
        streams[f"480x280_{skyname}_t{t}"] = {"pixels_sha256": U.sha(px), "stream_sha256": U.sha(buf[:nbytes]),
                                              "stream_bytes": int(nbytes), "skybox": skyname, "dim": dim, "t": t}
        print(skyname, nbytes, streams[f"480x280_{skyname}_t{t}"]["stream_sha256"][:16])
    # a non-default size through the sed-resized reference build
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "sized", "W=96", "H=56"])
    ref96 = U.load_reference(os.path.join(ROOT, "oracle", "_ref", "libtrt_ref_96x56.so"))
    sky = S.synthetic_cubemap("uv_gradient", 64)
    sc = S.SceneData(96, 56, sky).set_time(3.7)
    px = U.cpu_render(ref96, "project_scene", sc)
    buf = np.zeros(abi.stream_bytes(96, 56) + 64, dtype=np.uint8)
    nbytes = ref96.ref_draw_screen_bytes(C.byref(U.screen_for(px)), U.VP(buf.ctypes.data), buf.size)
    assert nbytes == abi.stream_bytes(96, 56)
    streams["96x56_uv_gradient_t3.7"] = {"pixels_sha256": U.sha(px), "stream_sha256": U.sha(buf[:nbytes]),
                                         "stream_bytes": int(nbytes), "skybox": "uv_gradient", "dim": 64, "t": 3.7}
    # encoder on arbitrary doubles (outside [0,1] too) at the reference's own size
    px = U.random_encoder_pixels()
    buf = np.zeros(abi.stream_bytes(480, 280) + 64, dtype=np.uint8)
    nbytes = ref.ref_draw_screen_bytes(C.byref(U.screen_for(px)), U.VP(buf.ctypes.data), buf.size)
    streams["480x280_random_pixels"] = {"pixels_sha256": U.sha(px), "stream_sha256": U.sha(buf[:nbytes]), "stream_bytes": int(nbytes)}
    with open(os.path.join(OUT, "streams.json"), "w") as f:
        json.dump(streams, f, indent=1, sort_keys=True)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
