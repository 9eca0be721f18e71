"""Parity at the sizes BASELINE.json puts the configs at, against the REFERENCE itself.

The checker is the reference's own project_scene in its row-range build (oracle/_ref/libtrt_ref_rows.so: the
unmodified TU with the row loop of TRT.c:973 bounded, oracle/Makefile), fanned over the host's cores; when that build
did not travel to the box the oracle port (pinned to the reference by tests/test_oracle.py) takes its place and the
report says so.  Compared bit for bit: the FP64 framebuffer (trt_project_scene, TRT.c:966-1069) and the terminal
stream (trt_render_ansi / trt_render_orbit, TRT.c:1142-1172).

  C1  demo scene 3840x2160, uv_checker, t = 3.7        full frame
  C2  demo scene 7680x4320, milky_way, t = 3.7         full frame
  C3  1024-sphere stress scene 3840x2160, uv_checker   32 evenly spaced rows + the rows either side of the 8-GPU band cuts
  C4  orbit 1920x1080, every 30th of the 360 frames    full frames, through trt_render_orbit

Every test writes the §8(d) mismatch list (row, col, channel, ref, got) and the identical-cell percentage into
gpurun_out/r02_parity_configs.json (expected: empty, 100 %).  north_star's tolerance is 1/255 per channel and 99.9 %
of the cells; enforced here: 0 and 100 %."""
import os
import time

import numpy as np
import pytest

from terminalraytracer_b200 import abi, scene as S, sharding
from tests import _util as U

pytestmark = pytest.mark.gpu

TOL_CHANNEL = 0.0
TOL_CELLS_PCT = 100.0


def _stream_rows(stream, w, h):
    return stream[abi.HOME_BYTES:-abi.TAIL_NULS].reshape(h, abi.row_bytes(w))


def _check(name, sc, rows, got_px_rows, got_stream_rows, orc, extra=None):
    t0 = time.perf_counter()
    want_px, kind = U.checker_rows_parallel(sc, rows, orc)
    cpu_s = time.perf_counter() - t0
    want_stream_rows = _stream_rows(U.oracle_stream(orc, want_px), sc.width, len(rows))
    rep = U.mismatch_report(got_px_rows, want_px, got_stream_rows, want_stream_rows, rows)
    rep.update({"config": name, "width": sc.width, "height": sc.height, "checker": kind,
                "checker_seconds": cpu_s, "checker_threads": os.cpu_count(),
                "checker_Mrays_per_s_all_cores": 10.0 * sc.width * len(rows) / cpu_s / 1e6,
                "tolerance_enforced": {"channel": TOL_CHANNEL, "cells_pct": TOL_CELLS_PCT}})
    if extra:
        rep.update(extra)
    U.write_report(name, rep)
    assert rep["max_abs_diff"] <= TOL_CHANNEL, rep["mismatches_row_col_channel_ref_got"][:10]
    assert rep["pixel_channel_mismatches"] == 0
    assert rep["cells_identical_pct"] >= TOL_CELLS_PCT and rep["row_terminators_identical"]
    return rep


@pytest.mark.parametrize("cfg", [("C1", "uv_checker", 3840, 2160), ("C2", "milky_way", 7680, 4320)], ids=lambda c: c[0])
def test_gpu_demo_scene_full_frame_vs_reference(renderer, orc, cfg):
    name, skyname, w, h = cfg
    sky = S.get_skybox(skyname)
    sc = S.SceneData(w, h, sky).set_time(3.7)
    renderer.upload_skybox(sky)
    got_px = renderer.project_scene(sc)                       # drop-in for project_scene: FP64 framebuffer
    stream = np.array(renderer.render_ansi(sc))               # K1 (quantised) -> K2 -> D2H
    assert stream.size == abi.stream_bytes(w, h)
    assert bytes(stream[:abi.HOME_BYTES]) == abi.HOME and not stream[-abi.TAIL_NULS:].any()
    _check(f"{name}_demo_{w}x{h}_{skyname}_t3.7", sc, range(h), got_px, _stream_rows(stream, w, h), orc,
           {"scope": "full frame", "skybox_source": "skybox/%s on disk" % skyname if os.path.isdir(os.path.join(U.ROOT, "skybox", skyname))
            else "synthetic stand-in (scene.synthetic_cubemap)"})


def test_gpu_stress_scene_rows_vs_reference(renderer, orc):
    """C3: 1024 spheres at 3840x2160 (k-d-sorted clusters, CULL == 2): 32 evenly spaced rows and the two rows either side of
    every equal-height 8-GPU band cut; brute force on the CPU is ~1.5 core-seconds per row"""
    w, h = 3840, 2160
    sky = S.get_skybox("uv_checker")
    sc = S.SceneData(w, h, sky, kind="stress", num_spheres=1024).set_time(3.7)
    renderer.upload_skybox(sky)
    rows = set(int(round(i * (h - 1) / 31.0)) for i in range(32))
    for r0, r1 in sharding.row_bands(h, 8):
        rows.update(r for r in (r0 - 1, r0, r1 - 1, r1) if 0 <= r < h)
    rows = sorted(rows)
    got_px = renderer.project_scene(sc)
    stream = np.array(renderer.render_ansi(sc))
    # the audit of the certificates on a slice of the same frame: every query answered both ways on the device
    renderer.set_scene(sc)
    ctr, _ = renderer.count_rows(w, h, h // 2, h // 2 + 16)
    assert ctr[28] == 0
    _check(f"C3_stress1024_{w}x{h}_uv_checker_t3.7", sc, rows, got_px[rows], _stream_rows(stream, w, h)[rows], orc,
           {"scope": "%d rows: 32 evenly spaced + the rows either side of the 8-GPU band cuts" % len(rows), "rows": rows,
            "device_audit_cull_violations_16_rows": int(ctr[28])})


def test_gpu_orbit_frames_vs_reference(renderer, orc):
    """C4: 360-frame orbit at 1920x1080: every 30th frame, full frame, delivered by trt_render_orbit (the streaming sink) and
    compared with the reference's pixels for that pose run through the encoder restatement"""
    w, h = 1920, 1080
    sky = S.get_skybox("milky_way")
    renderer.upload_skybox(sky)
    times = sharding.orbit_times(360)
    got = {}

    def sink(frame, view):
        got[frame] = np.array(view)
        return False

    assert renderer.render_orbit(S.SceneData(w, h, sky), times, sink, first=0, stride=30) == 12
    assert list(got) == list(range(0, 360, 30))
    total_bad = total_cells = 0
    for k in sorted(got):
        sc = S.SceneData(w, h, sky).set_time(times[k])
        px = renderer.project_scene(sc)
        rep = _check(f"C4_orbit_{w}x{h}_frame{k:03d}", sc, range(h), px, _stream_rows(got[k], w, h), orc,
                     {"scope": "full frame %d of 360 (t = %.6f s) through trt_render_orbit" % (k, times[k])})
        total_bad += rep["pixel_channel_mismatches"]
        total_cells += rep["cells_identical"]
    assert total_bad == 0 and total_cells == 12 * w * h
