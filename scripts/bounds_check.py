"""The substitute for compute-sanitizer on this GPU pool (the tool is closed by the operators: "runs under it have left GPUs needing a
reset"): libtrt_b200 built with -DTRT_BOUNDS_CHECK checks every computed index of its kernels on the device and counts violations
per site (csrc/trt_device.cuh TRT_BOUND, trt_debug_bounds).  This script builds that variant, runs the sanitizer workload
(scripts/sanitize_small.py: every kernel flavour, fused epilogue at four alignments, orbit sink, probes, 12 random scenes) and
full-size frames on it, and prints the counters — all must be 0.   usage: python scripts/bounds_check.py [out.txt]"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r02_bounds_check.txt")
lib = "libtrt_b200_checked.so"
if not os.path.exists(os.path.join(ROOT, "terminalraytracer_b200", lib)):
    subprocess.check_call([sys.executable, "-m", "terminalraytracer_b200.build", "--quiet", "--out=" + lib, "-DTRT_BOUNDS_CHECK"], cwd=ROOT)
env = dict(os.environ, TRT_B200_LIB=lib)
lines = ["libtrt_b200 built with -DTRT_BOUNDS_CHECK; compute-sanitizer itself is closed on this pool (see scripts/bounds_check.py)"]
r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "sanitize_small.py")], env=env, capture_output=True, text=True)
lines += ["== scripts/sanitize_small.py (exit %d)" % r.returncode] + r.stdout.strip().splitlines() + r.stderr.strip().splitlines()[-5:]
code = r"""
import ctypes as C, sys, numpy as np
sys.path.insert(0, %r)
from terminalraytracer_b200 import renderer as R, scene as S, sharding
sky = S.get_skybox("milky_way")
rd = R.Renderer(0, sky)
for (w, h, kind) in [(1920, 1080, "demo"), (3840, 2160, "demo"), (7680, 4320, "demo"), (1001, 333, "demo"), (480, 270, "stress")]:
    sc = S.SceneData(w, h, sky, kind=kind).set_time(3.7)
    n = np.array(rd.render_ansi(sc)).size
    print("rendered", kind, w, h, n, "bytes", flush=True)
got = []
rd.render_orbit(S.SceneData(640, 360, sky), sharding.orbit_times(24), lambda f, v: got.append(f) and False)
print("orbit frames", len(got), flush=True)
counts = (C.c_uint * 32)()
print("bounds checks compiled in:", bool(rd.L.trt_debug_bounds(counts)), " violations per site:", list(counts), flush=True)
rd.close()
""" % ROOT
r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
lines += ["== full-size frames (exit %d)" % r.returncode] + r.stdout.strip().splitlines() + r.stderr.strip().splitlines()[-5:]
open(out_path, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
