"""Randomised parity run: random scenes (sphere counts across the chunk / cluster limits, random radii and materials,
random lights above and below a randomly tilted ground, random camera poses incl. cameras inside the sphere cloud) through
the CUDA path, compared bit for bit with the oracle, plus the counting build's audit (every query answered both ways).
usage: fuzz_parity.py [scenes [seed]]   — prints one line per scene and a summary; exit code 1 on any mismatch."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from terminalraytracer_b200 import abi, renderer as R, scene as S
from tests import _util as U


def random_scene(rng, sky):
    n = int(rng.choice([0, 1, 2, 5, 6, 7, 8, 31, 32, 33, 63, 64, 65, 100, 257]))
    w, h = int(rng.integers(9, 70)), int(rng.integers(5, 40))
    sc = S.SceneData(w, h, sky, kind="stress", num_spheres=max(n, 1)) if n != 6 else S.SceneData(w, h, sky)
    sc.c.num_spheres = n
    spread = float(rng.choice([1.0, 3.0, 8.0]))
    for i in range(n):
        sp = sc.spheres[i]
        if n != 6:
            sp.center = abi.Vector(*(rng.uniform(-spread, spread, 3)))
            sp.radius = float(rng.uniform(0.05, 0.9))
        sp.material.reflectivity = float(rng.choice([0.0, 0.2, 0.8, 1.0]))
    nd, npnt = int(rng.integers(0, 4)), int(rng.integers(0, 4))
    sc.dls = (abi.DirectionalLight * max(nd, 1))()
    sc.pls = (abi.PointLight * max(npnt, 1))()
    for i in range(nd):
        sc.dls[i] = abi.DirectionalLight(abi.Vector(*rng.normal(size=3)), abi.Vector(*rng.uniform(0.1, 1, 3)))
    for i in range(npnt):
        sc.pls[i] = abi.PointLight(abi.Vector(*rng.uniform(-4, 4, 3)), abi.Vector(*rng.uniform(0.1, 1, 3)), float(rng.uniform(0.5, 30)))
    sc.c.directional_lights = C.cast(sc.dls, C.POINTER(abi.DirectionalLight))
    sc.c.num_directional_lights = nd
    sc.c.point_lights = C.cast(sc.pls, C.POINTER(abi.PointLight))
    sc.c.num_point_lights = npnt
    if rng.random() < 0.5:
        sc.c.ground.normal = abi.Vector(*(rng.normal(size=3) * 0.3 + np.array([0.0, 1.5, 0.0])))
        sc.c.ground.point = abi.Vector(0.0, float(rng.uniform(-3, -0.5)), 0.0)
    sc.set_time(float(rng.uniform(0, 20)))
    if rng.random() < 0.3:
        o = rng.uniform(-3, 3, 3)
        sc.c.camera.frame.origin = abi.Vector(float(o[0]), float(abs(o[1]) + 0.2), float(o[2]))
    return sc


def main():
    scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 2024
    rng = np.random.default_rng(seed)
    orc = U.load_oracle()
    sky = S.synthetic_cubemap("uv_gradient", 64)
    rd = R.Renderer(0, sky)
    bad = 0
    for k in range(scenes):
        sc = random_scene(rng, sky)
        got = rd.project_scene(sc)
        want = U.cpu_render(orc, "orc_project_scene", sc)
        rd.set_scene(sc)
        ctr, _ = rd.count_rows(sc.width, sc.height, 0, sc.height)
        same = np.array_equal(got, want)
        want_stream = U.oracle_stream(orc, want)
        stream_ok = np.array_equal(np.array(rd.render_ansi(sc)), want_stream)
        # the fused kernel (K1 encodes and stores its tiles), in two bands, at a random byte alignment of the stream
        shift = int(rng.integers(0, 4))
        d_stream = rd.L.trt_device_alloc(want_stream.size + 8)
        cut = int(rng.integers(0, sc.height + 1))
        rd.render_rows_ansi(sc.width, sc.height, 0, cut, d_stream + shift)
        rd.render_rows_ansi(sc.width, sc.height, cut, sc.height, d_stream + shift)
        rd.stream_frame(d_stream + shift, sc.width, sc.height)
        fused = np.empty(want_stream.size, dtype=np.uint8)
        rd.L.trt_copy_to_host(fused.ctypes.data_as(C.c_void_p), C.c_void_p(d_stream + shift), want_stream.size)
        rd.synchronize()
        rd.L.trt_device_free(d_stream)
        stream_ok = stream_ok and np.array_equal(fused, want_stream)
        if not (same and stream_ok and ctr[28] == 0):
            bad += 1
        print("scene %3d: %3dx%-3d spheres %3d lights %d+%d  pixels %s  stream %s  audit disagreements %d" % (
            k, sc.width, sc.height, sc.c.num_spheres, sc.c.num_directional_lights, sc.c.num_point_lights,
            "==" if same else "DIFFER (max %.3g)" % np.abs(got - want).max(), "==" if stream_ok else "DIFFERS", ctr[28]), flush=True)
    rd.close()
    print("fuzz: %d scenes, %d with a mismatch" % (scenes, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
