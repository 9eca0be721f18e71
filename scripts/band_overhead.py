"""Development aid (one GPU): what row-band sharding costs K1 — the frame as ONE launch against the same frame as N equal-cost
band launches (the bands bench.py would give N ranks), every launch CUDA-event timed; also the fused-epilogue flavour."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from terminalraytracer_b200 import abi, renderer as R, scene as S, sharding
w, h = 7680, 4320
sky = S.get_skybox("milky_way")
rd = R.Renderer(0, sky)
rd.use_stream(torch.cuda.current_stream().cuda_stream)
sc = S.SceneData(w, h, sky).set_time(3.7)
weights = rd.estimate_row_costs(sc)
rd.set_scene(sc)
quant = torch.zeros(w * h * 4, dtype=torch.uint8, device="cuda")
stream = torch.zeros(abi.stream_bytes(w, h) + 64, dtype=torch.uint8, device="cuda")


def timed(fn, reps=5):
    best = 1e9
    for i in range(reps + 1):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i:
            best = min(best, a.elapsed_time(b))
    return best


full = timed(lambda: rd.render_rows_quant(w, h, 0, h, quant.data_ptr()))
print("one launch: %.3f ms" % full, flush=True)
for n in (2, 4, 8):
    bands = sharding.row_bands(h, n, weights)
    t = [timed(lambda r0=r0, r1=r1: rd.render_rows_quant(w, h, r0, r1, quant.data_ptr() + r0 * w * 4)) for (r0, r1) in bands]
    tf = [timed(lambda r0=r0, r1=r1: rd.render_rows_ansi(w, h, r0, r1, stream.data_ptr())) for (r0, r1) in bands]
    print("N=%d bands: sum %.3f ms (+%.1f %%), max %.3f, ideal %.3f | fused epilogue: sum %.3f (+%.1f %%), max %.3f  | rows %s" % (
        n, sum(t), 100 * (sum(t) / full - 1), max(t), full / n, sum(tf), 100 * (sum(tf) / full - 1), max(tf), [b[1] - b[0] for b in bands]), flush=True)
rd.use_stream(None)
rd.close()
