"""Development aid: how well do the cost-weighted row bands balance K1?  Renders every band of the bench frame on ONE GPU
and prints max/mean of the per-band K1 times for 2, 4 and 8 ranks (what the slowest rank would cost in a sharded step)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from terminalraytracer_b200 import renderer as R, scene as S, sharding
w, h = 7680, 4320
sky = S.get_skybox("milky_way")
rd = R.Renderer(0, sky)
sc = S.SceneData(w, h, sky).set_time(3.7)
costs = rd.estimate_row_costs(sc)
rd.use_stream(torch.cuda.current_stream().cuda_stream)
rd.set_scene(sc)
q = torch.empty(w * h * 4, dtype=torch.uint8, device="cuda")
for world in (2, 4, 8):
    for name, weights in (("weighted", costs), ("equal rows", None)):
        bands = sharding.row_bands(h, world, weights)
        ms = []
        for (r0, r1) in bands:
            best = 1e9
            for _ in range(2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rd.render_rows_quant(w, h, r0, r1, q.data_ptr())
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            ms.append(best)
        print("N=%d %-10s max %.2f ms mean %.2f ms  max/mean %.3f  bands %s" % (world, name, max(ms), sum(ms) / len(ms), max(ms) / (sum(ms) / len(ms)), [b[1] - b[0] for b in bands]))
rd.close()
