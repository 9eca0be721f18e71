"""Shared helpers for the test-suite: loading the checkers (oracle restatement, reference build) and
calling them on the same ctypes structures the product receives.  Test infrastructure only."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

from terminalraytracer_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libtrt_oracle.so")
CERTCHECK_SO = os.path.join(ROOT, "oracle", "_build", "libtrt_certcheck.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libtrt_ref.so")
REF_ROWS_SO = os.path.join(ROOT, "oracle", "_ref", "libtrt_ref_rows.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE_DIR = "/root/reference"

VP = C.c_void_p


class Counters(C.Structure):
    _fields_ = [(n, C.c_longlong) for n in (
        "sphere_tests", "sphere_disc_ok", "sphere_t0_pos", "sphere_closest", "plane_tests", "plane_denom_ok",
        "plane_t_pos", "plane_closest", "sky_lookups", "trace_calls", "trace_hits", "lighting_calls",
        "bounce_iters", "samples", "pixels")] + [("bounce_hist", C.c_longlong * (abi.BOUNCE_LIMIT + 1))]

    def as_gpu_order(self):
        """same order as CounterId in csrc/trt_device.cuh"""
        head = [self.sphere_tests, self.sphere_disc_ok, self.sphere_t0_pos, self.sphere_closest, self.plane_tests,
                self.plane_denom_ok, self.plane_t_pos, self.plane_closest, self.sky_lookups, self.trace_calls,
                self.trace_hits, self.lighting_calls, self.bounce_iters, self.samples, self.pixels]
        return head + list(self.bounce_hist)


def build_oracle():
    if not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return ORACLE_SO


def load_oracle():
    lib = C.CDLL(build_oracle())
    lib.orc_project_scene.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.Screen)]
    lib.orc_render_rows.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.Screen), C.c_int, C.c_int, C.POINTER(Counters)]
    lib.orc_encode_stream.argtypes = [C.POINTER(abi.Screen), VP]
    lib.orc_encode_stream.restype = C.c_size_t
    lib.orc_encode_rows.argtypes = [C.POINTER(abi.Screen), C.c_int, C.c_int, VP]
    lib.orc_encode_rows.restype = C.c_size_t
    lib.orc_stream_bytes.argtypes = [C.c_int, C.c_int]
    lib.orc_stream_bytes.restype = C.c_size_t
    lib.orc_hit_sphere.argtypes = [C.POINTER(abi.Ray), C.POINTER(abi.Sphere), C.POINTER(abi.Vector), VP]
    lib.orc_hit_plane.argtypes = [C.POINTER(abi.Ray), C.POINTER(abi.Plane), C.POINTER(abi.Vector), VP]
    lib.orc_sky_texel.argtypes = [C.POINTER(abi.Skybox), C.POINTER(abi.Vector), C.POINTER(abi.Color), C.POINTER(C.c_int),
                                  C.POINTER(C.c_long)]
    lib.orc_closest_hit.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.Ray), C.POINTER(abi.Vector), C.POINTER(abi.Vector),
                                    C.POINTER(abi.Material), VP]
    lib.orc_closest_hit.restype = C.c_int
    lib.orc_light_surface.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.Vector), C.POINTER(abi.Vector),
                                      C.POINTER(abi.Material), VP]
    lib.orc_subpixel_offsets.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.orc_model_flops.argtypes = [C.POINTER(Counters)]
    lib.orc_model_flops.restype = C.c_double
    lib.orc_sizeof_counters.restype = C.c_size_t
    return lib


CERT_STATS = ("primary", "primary_tile_survivors", "primary_ground_culled", "bounce", "bounce_survivors", "bounce_ground_culled",
              "bounce_exact_hits", "dir", "dir_open", "dir_blocked", "dir_unknown", "point", "point_open", "point_blocked",
              "point_unknown", "shadow_exact_tests", "sky", "sky_certified", "cluster_tests", "clusters_missed", "subcluster_tests", "subclusters_missed", "tiles", "tiles_patch", "tiles_patch_empty", "patch_dir_candidates", "patch_point_candidates",
              "patch_bounce_candidates", "patch_records")


def load_certcheck():
    """oracle/cert_check.c: the product's float certificates (csrc/trt_cert.h) audited on the CPU against the oracle"""
    if not os.path.exists(CERTCHECK_SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    lib = C.CDLL(CERTCHECK_SO)
    lib.cert_check_rows.restype = C.c_longlong
    lib.cert_check_rows.argtypes = [C.POINTER(abi.Scene), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
    assert lib.cert_check_num_stats() == len(CERT_STATS)
    return lib


def cert_check(lib, scene, row0=0, row1=None):
    st = (C.c_longlong * len(CERT_STATS))()
    bad = lib.cert_check_rows(C.byref(scene.c), scene.width, scene.height, row0, scene.height if row1 is None else row1, st)
    return bad, dict(zip(CERT_STATS, list(st)))


def have_reference_build():
    return os.path.exists(REF_SO)


def load_reference(path=REF_SO):
    lib = C.CDLL(path)
    lib.project_scene.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.Screen)]
    lib.ref_orbit_camera.argtypes = [C.POINTER(abi.Camera), C.c_double]
    lib.ref_subpixel_offsets.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.ref_draw_screen_bytes.argtypes = [C.POINTER(abi.Screen), VP, C.c_long]
    lib.ref_draw_screen_bytes.restype = C.c_long
    lib.ref_time_project_scene.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.Screen)]
    lib.ref_time_project_scene.restype = C.c_double
    lib.ray_intersects_sphere.argtypes = [C.POINTER(abi.Ray), C.POINTER(abi.Sphere), C.POINTER(abi.Vector)]
    lib.ray_intersects_plane.argtypes = [C.POINTER(abi.Ray), C.POINTER(abi.Plane), C.POINTER(abi.Vector)]
    lib.get_skybox_color.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.Vector), C.POINTER(abi.Color)]
    lib.trace_ray.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.Ray), C.POINTER(abi.Vector), C.POINTER(abi.Vector),
                              C.POINTER(abi.Material)]
    lib.trace_ray.restype = C.c_int
    lib.apply_lighting.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.Vector), C.POINTER(abi.Vector), C.POINTER(abi.Vector),
                                   C.POINTER(abi.Material)]
    lib.byte_to_digits.argtypes = [C.c_int, C.c_char_p]
    for name in ("ref_sizeof_scene", "ref_sizeof_sphere", "ref_sizeof_plane", "ref_sizeof_camera", "ref_sizeof_skybox",
                 "ref_sizeof_screen", "ref_offsetof_scene_camera", "ref_offsetof_scene_skybox", "ref_offsetof_scene_ground",
                 "ref_offsetof_scene_point_lights", "ref_screenbuffer_bytes"):
        getattr(lib, name).restype = C.c_size_t
    return lib


def screen_for(px):
    h, w, _ = px.shape
    return abi.Screen(px.ctypes.data_as(C.POINTER(abi.Vector)), w, h)


def cpu_render(lib, fn_name, scene):
    """Run a CPU checker's project_scene-like function on a SceneData; returns (H,W,3) float64."""
    px = np.zeros((scene.height, scene.width, 3), dtype=np.float64)
    scr = screen_for(px)
    getattr(lib, fn_name)(C.byref(scene.c), C.byref(scr))
    return px


def oracle_rows(orc, scene, row0, row1, counters=None):
    px = np.zeros((scene.height, scene.width, 3), dtype=np.float64)
    scr = screen_for(px)
    orc.orc_render_rows(C.byref(scene.c), C.byref(scr), row0, row1, C.byref(counters) if counters is not None else None)
    return px[row0:row1]


def oracle_stream(orc, px):
    h, w, _ = px.shape
    px = np.ascontiguousarray(px, dtype=np.float64)
    out = np.zeros(abi.stream_bytes(w, h), dtype=np.uint8)
    n = orc.orc_encode_stream(C.byref(screen_for(px)), VP(out.ctypes.data))
    assert n == out.size
    return out


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def hash_uniform(shape, seed, lo=0.0, hi=1.0):
    """Deterministic uniform doubles from splitmix64 (integer arithmetic only, so fixtures regenerate
    identically on any numpy version)."""
    from terminalraytracer_b200.scene import _splitmix64
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        z = _splitmix64(np.arange(n, dtype=np.uint64) + (np.uint64(seed) << np.uint64(32)))
    u = (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return (lo + u * (hi - lo)).reshape(shape)


def random_encoder_pixels():
    """The 480x280 input of the 'random pixels' encoder golden (values outside [0,1] included)."""
    px = hash_uniform((280, 480, 3), 77, -0.2, 1.3)
    px[0, :16, 0] = np.arange(16) / 255.0
    return px


# ---- checkers fanned over the host's cores (full-size configs) ------------------------------------------

def load_reference_rows():
    """oracle/_ref/libtrt_ref_rows.so: the reference TU with the row loop of TRT.c:973 bounded by a thread-local
    [row0,row1) (oracle/Makefile); None when the reference build did not travel."""
    if not os.path.exists(REF_ROWS_SO):
        return None
    lib = C.CDLL(REF_ROWS_SO)
    lib.ref_project_rows.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.Screen), C.c_int, C.c_int]
    return lib


def checker_rows_parallel(scene, rows, orc=None, threads=None):
    """FP64 pixels of the given rows (any iterable of row indices) of scene.width x scene.height as the REFERENCE's
    project_scene computes them (row-range build), or as the oracle port does when the reference build is absent;
    the rows are spread over the host cores (ctypes releases the GIL; the row bounds are thread-local).
    Returns (array (len(rows), W, 3), kind)."""
    from concurrent.futures import ThreadPoolExecutor
    rows = [int(r) for r in rows]
    w, h = scene.width, scene.height
    ref = load_reference_rows()
    kind = "reference" if ref is not None else "port"
    if ref is None and orc is None:
        orc = load_oracle()
    threads = threads or max(1, (os.cpu_count() or 1))
    out = np.zeros((len(rows), w, 3), dtype=np.float64)
    # contiguous runs of rows -> one call each, cut so that every thread gets several jobs (rows differ ~5x in cost)
    runs, start = [], 0
    for i in range(1, len(rows) + 1):
        if i == len(rows) or rows[i] != rows[i - 1] + 1:
            runs.append((start, i))
            start = i
    jobs = []
    max_run = max(1, len(rows) // (threads * 8))
    for a, b in runs:
        for s in range(a, b, max_run):
            jobs.append((s, min(b, s + max_run)))

    def work(job):
        a, b = job
        r0, r1 = rows[a], rows[b - 1] + 1
        # a Screen whose row r0 lands at out[a]: the checkers index pixels[row * width + col]
        base = out[a:].ctypes.data - r0 * w * 24
        scr = abi.Screen(C.cast(C.c_void_p(base), C.POINTER(abi.Vector)), w, h)
        if ref is not None:
            ref.ref_project_rows(C.byref(scene.c), C.byref(scr), r0, r1)
        else:
            orc.orc_render_rows(C.byref(scene.c), C.byref(scr), r0, r1, None)

    with ThreadPoolExecutor(threads) as pool:
        list(pool.map(work, jobs))
    return out, kind


def mismatch_report(got_px, want_px, got_stream_rows, want_stream_rows, row_ids, limit=200):
    """mismatches as (row, col, channel, ref, got) for the FP64 pixels and the identical-cell fraction of
    the terminal stream (rows of 25-byte cells + '\\n').  got/want_px: (R, W, 3); *_stream_rows: (R, 25W+1) uint8."""
    bad = np.argwhere(got_px != want_px)
    listed = [(int(row_ids[r]), int(c), int(ch), float(want_px[r, c, ch]), float(got_px[r, c, ch])) for r, c, ch in bad[:limit]]
    R, W = got_px.shape[0], got_px.shape[1]
    gc = got_stream_rows[:, :-1].reshape(R, W, abi.CELL_BYTES)
    wc = want_stream_rows[:, :-1].reshape(R, W, abi.CELL_BYTES)
    same_cells = int((gc == wc).all(axis=2).sum())
    newline_ok = bool((got_stream_rows[:, -1] == want_stream_rows[:, -1]).all())
    return {"rows_checked": int(R), "pixels_checked": int(R * W), "pixel_channel_mismatches": int(len(bad)),
            "max_abs_diff": float(np.abs(got_px - want_px).max()) if got_px.size else 0.0,
            "mismatches_row_col_channel_ref_got": listed,
            "cells_checked": int(R * W), "cells_identical": same_cells, "cells_identical_pct": 100.0 * same_cells / max(R * W, 1),
            "row_terminators_identical": newline_ok}


def write_report(name, report):
    """per-config parity reports: gpurun_out/ on the GPU box (merged back by gpurun), copied to profiles/ by hand"""
    import json
    d = os.environ.get("TRT_REPORT_DIR", os.path.join(ROOT, "gpurun_out"))
    try:
        os.makedirs(d, exist_ok=True)
        path = os.path.join(d, "r02_parity_configs.json")
        try:
            with open(path) as f:
                allr = json.load(f)
        except (OSError, ValueError):
            allr = {}
        allr[name] = report
        with open(path, "w") as f:
            json.dump(allr, f, indent=1)
    except OSError:
        pass
