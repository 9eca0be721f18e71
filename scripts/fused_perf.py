"""Development aid: K1 + K2 against the fused K1 (trt_render_rows_ansi_device) writing into device memory and into page-locked
host memory, one GPU, 7680x4320 demo frame."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from terminalraytracer_b200 import abi, pipeline, renderer as R, scene as S
w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (7680, 4320)
sky = S.get_skybox("milky_way")
rd = R.Renderer(0, sky)
sc = S.SceneData(w, h, sky).set_time(3.7)
rd.use_stream(torch.cuda.current_stream().cuda_stream)
rd.set_scene(sc)
total = abi.stream_bytes(w, h)
quant = torch.empty(w * h * 4, dtype=torch.uint8, device="cuda")
dev = torch.empty(total + 16, dtype=torch.uint8, device="cuda")
shared = pipeline.SharedHostStream(rd, total)


def timed(fn, n=3):
    best = 1e30
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def two_kernels():
    rd.render_rows_quant(w, h, 0, h, quant.data_ptr())
    rd.encode_rows_quant(quant.data_ptr(), w, h, dev.data_ptr(), abi.HOME_BYTES)


print("K1 + K2 (device)        %8.3f ms" % timed(two_kernels))
print("fused K1 -> device      %8.3f ms" % timed(lambda: rd.render_rows_ansi(w, h, 0, h, dev.data_ptr())))
print("fused K1 -> pinned host %8.3f ms" % timed(lambda: rd.render_rows_ansi(w, h, 0, h, shared.ptr)))
rd.use_stream(None)
t = []
for _ in range(3):
    t0 = time.perf_counter(); rd.render_ansi(sc); t.append(time.perf_counter() - t0)
print("trt_render_ansi e2e     %8.3f ms" % (min(t) * 1e3))
shared.close()
rd.close()
