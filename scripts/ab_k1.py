"""Development aid: A/B of experiment builds of libtrt_b200 (python -m terminalraytracer_b200.build --out=libtrt_b200_x.so -D...).
    python scripts/ab_k1.py libtrt_b200.so libtrt_b200_x.so ...
For every library (one subprocess each): K1 alone, one launch per 7680x4320 frame (quantised cells), CUDA-event timed on the
launching stream, best and median of 7 after 2 warm-ups, sha256 of the cells (must agree across builds), and the 1024-sphere
stress scene at 1920x1080."""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def measure():
    sys.path.insert(0, ROOT)
    import torch
    from terminalraytracer_b200 import renderer as R, scene as S
    sky = S.get_skybox("milky_way")
    rd = R.Renderer(0, sky)
    rd.use_stream(torch.cuda.current_stream().cuda_stream)
    out = {}
    for (w, h, kind, reps) in [(7680, 4320, "demo", 7), (1920, 1080, "demo", 5), (1920, 1080, "stress", 2)]:
        sc = S.SceneData(w, h, sky, kind=kind).set_time(3.7)
        rd.set_scene(sc)
        quant = torch.zeros(w * h * 4, dtype=torch.uint8, device="cuda")
        times = []
        for i in range(reps + 2):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rd.render_rows_quant(w, h, 0, h, quant.data_ptr())
            b.record()
            torch.cuda.synchronize()
            if i >= 2:
                times.append(a.elapsed_time(b))
        times.sort()
        out["%s_%dx%d" % (kind, w, h)] = {"best_ms": times[0], "median_ms": times[len(times) // 2],
                                          "sha": hashlib.sha256(quant.cpu().numpy().tobytes()).hexdigest()[:16]}
    rd.use_stream(None)
    rd.close()
    print("AB_RESULT " + json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) == 1:
        measure()
        sys.exit(0)
    base = None
    for lib in sys.argv[1:]:
        env = dict(os.environ, TRT_B200_LIB=lib)
        r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("AB_RESULT ")]
        if not line:
            print(lib, "FAILED", r.stdout[-500:], r.stderr[-1500:])
            continue
        res = json.loads(line[0][len("AB_RESULT "):])
        if base is None:
            base = res
        print("%-34s" % lib + "  ".join("%s %.3f/%.3f ms %s" % (k, v["best_ms"], v["median_ms"], "same" if v["sha"] == base[k]["sha"] else "DIFFERENT")
                                        for k, v in res.items()), flush=True)
