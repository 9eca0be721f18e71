"""CPU audit of the single-precision certificates the render kernel uses (csrc/trt_cert.h).

oracle/cert_check.c compiles the SAME header the kernel compiles and walks the reference's render loop through
the oracle; every decision the kernel would take from a certificate (sphere cannot be hit, light certainly
blocked, ground cannot matter, tile-level culls of primary rays) — and every decision it would take from exact
tests on the certificates' survivors only — is compared with the exact decision over all objects.  The bar is 0
contradictions; the statistics show that the certificates actually remove the work."""
import ctypes as C

import numpy as np
import pytest

from terminalraytracer_b200 import abi, scene as S
from tests import _util as U


@pytest.fixture(scope="module")
def cc():
    return U.load_certcheck()


SKY = S.synthetic_cubemap("uv_gradient", 64)


@pytest.mark.parametrize("t", [0.0, 3.7, 8.1, 13.3], ids=lambda t: f"t{t}")
def test_certificates_demo_scene_orbit(cc, t):
    bad, st = U.cert_check(cc, S.SceneData(240, 136, SKY).set_time(t))
    assert bad == 0
    assert st["primary"] == 240 * 136 * 10
    # the certificates decide almost everything: < 0.2 exact sphere tests per bounce ray (6 spheres) ...
    assert st["bounce_survivors"] < 0.2 * st["bounce"]
    # ... every surviving sphere of a bounce ray is (nearly) a real hit ...
    assert st["bounce_survivors"] <= st["bounce_exact_hits"] * 1.05 + 100
    # ... directional shadows are decided in float >= 95% of the time, and shadow queries need < 0.5 exact tests each
    assert st["dir_unknown"] < 0.05 * st["dir"]
    assert st["shadow_exact_tests"] < 0.5 * (st["dir"] + st["point"])


def test_certificates_stress_scene(cc):
    sc = S.SceneData(64, 36, SKY, kind="stress", num_spheres=1024).set_time(3.7)
    bad, st = U.cert_check(cc, sc)
    assert bad == 0
    assert st["primary_tile_survivors"] < 0.05 * 1024 * st["primary"]
    assert st["bounce_survivors"] < 0.005 * 1024 * st["bounce"]
    # k-d-sorted clusters of 32 (and of 8 inside them) with bounding balls: most are certainly missed by a given ray
    assert st["cluster_tests"] > 0 and st["clusters_missed"] > 0.4 * st["cluster_tests"]
    assert st["subclusters_missed"] > 0.5 * st["subcluster_tests"]


def test_certificates_far_camera_and_camera_inside_sphere(cc):
    far = S.SceneData(96, 54, SKY)
    far.c.camera.frame.origin = abi.Vector(0.0, 300.0, 4000.0)     # huge coordinates near the horizon
    inside = S.SceneData(96, 54, SKY)
    inside.c.camera.frame.origin = abi.Vector(1.0, 0.1, 0.0)       # eye inside sphere 0
    for sc in (far, inside):
        bad, st = U.cert_check(cc, sc)
        assert bad == 0
        assert st["bounce_survivors"] < 0.5 * st["bounce"]


def test_certificates_random_lights_tilted_ground_scaled_scene(cc):
    """lights in general position (also below the ground / inside spheres), an un-normalised tilted plane, and the
    whole scene scaled by 1e-3 and 1e+3 (the slacks are relative to the scene's magnitude)"""
    rng = np.random.default_rng(11)
    for trial in range(6):
        sc = S.SceneData(64, 36, SKY).set_time(float(rng.uniform(0, 20)))
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
        if trial % 2:
            sc.c.ground.normal = abi.Vector(0.1, 2.0, -0.3)
            sc.c.ground.point = abi.Vector(0.0, -1.5, 0.0)
        scale = (1.0, 1e-3, 1e3)[trial % 3]
        if scale != 1.0:
            for i in range(sc.c.num_spheres):
                sp = sc.spheres[i]
                sp.center = abi.Vector(sp.center.x * scale, sp.center.y * scale, sp.center.z * scale)
                sp.radius *= scale
            o = sc.c.camera.frame.origin
            sc.c.camera.frame.origin = abi.Vector(o.x * scale, o.y * scale, o.z * scale)
            gp = sc.c.ground.point
            sc.c.ground.point = abi.Vector(gp.x * scale, gp.y * scale, gp.z * scale)
            for i in range(3):
                q = sc.pls[i].position
                sc.pls[i].position = abi.Vector(q.x * scale, q.y * scale, q.z * scale)
        bad, _ = U.cert_check(cc, sc)
        assert bad == 0, (trial, scale)


def test_work_deal_of_a_chunk_covers_every_item_once():
    """classify_chunk (trt_render.cu) deals the (ray, group of 8) items of a chunk to quads of the lanes that run the query.  A
    restatement of its index arithmetic — the work list by group then ray, quads over the RANKS of the active lanes, a short
    "quad" when fewer than four lanes are active, the pair loop cut at the chunk's sphere count — for arbitrary lane masks:
    every pair of every needed group is classified by exactly one lane, for exactly its ray, and nothing else is touched."""
    rng = np.random.default_rng(11)
    masks = [0xffffffff, 0x1, 0x80000000, 0x7, 0xf, 0x1f, 0xaaaaaaaa, 0x0000ffff] + [int(rng.integers(1, 1 << 32)) for _ in range(40)]
    for lanes in masks:
        active = [l for l in range(32) if (lanes >> l) & 1]
        n_act = len(active)
        rank = {l: i for i, l in enumerate(active)}
        cnt = int(rng.choice([32, 31, 24, 17, 9, 8, 2, 1]))
        groups = {l: int(rng.integers(0, 16)) for l in active}
        # the work list: four ballots, position = items of the lower groups + lower lanes with the same group
        items = []
        for g in range(4):
            items += [(l << 2) | g for l in active if (groups[l] >> g) & 1]
        assert len(items) <= 128 and all(0 <= it < 128 for it in items)
        full = n_act >= 4
        width, quads = (4, n_act >> 2) if full else (n_act, 1)
        seen = {}
        for l in active:
            quad, first = (rank[l] >> 2, rank[l] & 3) if full else (0, rank[l])
            if quad >= quads:
                continue
            for it in range(quad, len(items), quads):
                r, g = items[it] >> 2, items[it] & 3
                for pp in range(first, 4, width):
                    j = 8 * g + 2 * pp
                    if j >= cnt:
                        break
                    assert (r, j) not in seen
                    seen[(r, j)] = l
        want = {(l, 8 * g + 2 * pp) for l in active for g in range(4) if (groups[l] >> g) & 1 for pp in range(4) if 8 * g + 2 * pp < cnt}
        assert set(seen) == want
