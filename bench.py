#!/usr/bin/env python
"""bench.py — headline benchmark of the render path (BASELINE.json: Mrays/s on the default demo scene
at 7680x4320, reference bounce depth, milky_way skybox, row-band sharded over 1/2/4/8 GPUs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one frame through the hot path: K1 (render band, persistent FP64 kernel) -> K2 (ANSI
encode) -> gather of the byte bands on rank 0 over NVLink peer memory (piece by piece behind the next
piece's K1, or — 8 GPUs — by K1 itself, which then encodes and stores its tiles straight into rank 0's
stream); the bands follow the ranks' measured K1 times.  Prints ONE JSON line on rank 0 (stdout carries
nothing else).

  value      Mrays/s = 10*W*H*K / seconds (primary samples; TRT.c:58 RAYS_PER_PIXEL = 10), scene and
             skybox resident in HBM, device-timed (CUDA events), max over ranks.
  e2e        same metric through the C-ABI call a host program makes (trt_render_ansi at N=1: scene
             upload H2D, K1, K2, D2H of the byte stream into pinned host memory — all inside the timing;
             at N>1 every rank copies its own bands into one shared page-locked host stream over its own
             PCIe link, pipeline.FramePipeline(host_stream=...)).
  roofline   K1 against the FP32 CUDA-core peak (the path has no dense contraction and ~400 flop/byte,
             SURVEY.md §8d): achieved = algorithmic flops of the frame (work counters x the as-written
             per-event flop costs, counted by the kernel itself in a separate untimed launch and
             cross-checked against the oracle's counters in tests) / K1's CUDA-event time.
             peak = FFMA throughput measured on this GPU by libtrt_b200 (MEASURED_PEAKS.json has no
             CUDA-core figure).  The encoder's HBM roofline is reported beside it.
  cpu_baseline  the UNMODIFIED reference project_scene (oracle/_ref, built from /root/reference by
             oracle/Makefile) on one host core, on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT, SKYBOX, T_POSE = 7680, 4320, "milky_way", 3.7
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` capture of this workload on one GPU
# (profiles/r01f_k1_k2_ncu_summary.txt): K1 15.4 MB + 107.9 MB (the 133 MB of quantised cells, partly still in L2 at the end;
# scene, skybox and the sample scratch stay in L2), K2 132.7 MB + 770.6 MB (the rest of the 829 MB stream was still in L2)
NCU_DRAM_BYTES = {"k_render": 15.381e6 + 107.915e6, "k_encode": 132.727e6 + 770.604e6}
CPU_SAMPLE_W, CPU_SAMPLE_H = 480, 270   # same 16:9 framing, 1/256 of the pixels


def workload_config(n_gpus):
    return {
        "workload": f"default demo scene (6 spheres + checker ground + 1 directional + 1 point light), {WIDTH}x{HEIGHT} cells, "
                    f"10 samples/pixel, bounce limit 10, skybox {SKYBOX} (synthetic 1024^2 x 6 stand-in: the reference's "
                    f"milky_way assets are not in its checkout), orbit pose t={T_POSE}s",
        "width": WIDTH, "height": HEIGHT, "samples_per_pixel": 10, "bounce_limit": 10, "skybox": SKYBOX,
        "sharding": (f"cost-weighted contiguous row-bands x{n_gpus} (1/8-resolution cost pre-pass, then feedback from the ranks' measured K1 "
                     f"times); every rank's encoded bytes go into rank 0's stream over NVLink peer memory (see 'gather'); one small NCCL "
                     f"all-gather (the K1 times) ends the step") if n_gpus > 1 else "single GPU",
        "l2": "no explicit flush: each step writes 133 MB of cells + 829 MB of stream (> 126 MB L2); inputs (scene 1 KB, "
              "skybox 25 MB) are meant to stay cache resident, the kernel is ALU-bound",
    }


# --------------------------------------------------------------------------------------------------------
# clocks sampled DURING the timed region

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thread = index, [], None, None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for stamp, r in self.rows:
            if len(r) < 7 or (self.t0 and stamp < self.t0) or (self.t1 and stamp > self.t1 + 0.05):
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); power.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------------
# CPU reference timing (the only place bench.py executes anything under oracle/)

def time_reference_cpu(width, height, repeats):
    """project_scene of the unmodified reference (oracle/_ref) — or of the oracle port when the reference
    build is absent — on ONE host thread (the reference is single-threaded as written)."""
    import numpy as np
    from terminalraytracer_b200 import abi, scene as S
    from tests import _util as U
    if U.have_reference_build():
        lib, fn, kind = U.load_reference(), "project_scene", "reference"
    else:
        lib, fn, kind = U.load_oracle(), "orc_project_scene", "port"
    sky = S.get_skybox(SKYBOX)
    sc = S.SceneData(width, height, sky).set_time(T_POSE)
    px = np.zeros((height, width, 3))
    scr = abi.Screen(px.ctypes.data_as(C.POINTER(abi.Vector)), width, height)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        getattr(lib, fn)(C.byref(sc.c), C.byref(scr))
        times.append(time.perf_counter() - t0)
    return kind, times


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    kind, times = time_reference_cpu(CPU_SAMPLE_W, CPU_SAMPLE_H, args.warmup + args.steps)
    timed = times[args.warmup:]
    total = sum(timed)
    rays = 10.0 * CPU_SAMPLE_W * CPU_SAMPLE_H * len(timed)
    value = rays / total / 1e6
    sample = (f"{len(timed)} frames of the same scene, pose and skybox at {CPU_SAMPLE_W}x{CPU_SAMPLE_H} "
              f"(1/256 of the pixels of the {WIDTH}x{HEIGHT} workload; cost per ray is resolution independent), "
              f"1 thread as written, {'unmodified reference TU' if kind == 'reference' else 'oracle port'} built -O3 -ffp-contract=off")
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": 1, "kind": kind, "sample": sample,
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--height", type=int, default=HEIGHT)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fused", type=int, default=int(os.environ.get("TRT_BENCH_FUSED", "-1")),
                    help="N > 1, device-resident gather: 1 = K1 stores the encoded tiles straight into rank 0's stream over NVLink (no K2, "
                         "no copies), 0 = K1, K2 and copy-engine pushes piece by piece, -1 = by GPU count (measured: pieces win at 2 GPUs, "
                         "14.8 vs 15.7 ms, the fused kernel at 8, 4.21 vs 4.38 ms)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    # stdout carries exactly one JSON line: whatever libraries print there (NCCL's version banner) goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import torch.distributed as dist
    from terminalraytracer_b200 import abi, pipeline, renderer as R, scene as S

    width, height = args.width, args.height
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the render path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    rd = R.Renderer(local_rank)
    sky = S.get_skybox(SKYBOX)
    rd.upload_skybox(sky)
    sc = S.SceneData(width, height, sky).set_time(T_POSE)
    # cost-weighted row bands (sky rows are ~5x cheaper than sphere/ground rows): every rank runs the same
    # deterministic 1/8-resolution pre-pass and derives the same bands; untimed, once per scene
    weights = rd.estimate_row_costs(sc) if world > 1 else None
    fused = (world >= 8) if args.fused < 0 else bool(args.fused)
    # N > 1: every rank pushes its encoded pieces into rank 0's stream over NVLink peer memory while its next piece renders
    # (or, fused: K1 itself stores every finished tile's bytes there);
    # the collective that ends a step carries the ranks' K1 times and the next step's bands follow from them (adapt)
    pipe = pipeline.FramePipeline(rd, width, height, rank, world, row_weights=weights, peer=world > 1, pieces=(0.7, 0.3), adapt=world > 1, fused=fused)
    stream = torch.cuda.current_stream()

    peaks = rd.measure_peaks() if rank == 0 else None

    # ---- device-resident steps ------------------------------------------------------------------------
    k1_events = []

    def step():
        pipe.render_local(sc, k1_events)     # set_scene, then per piece: K1, K2 (and the push to rank 0)
        pipe.gather()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()          # nvidia-smi takes a while to start: launch it before the warm-up,
    for _ in range(args.warmup):  # keep only the samples that fall inside the timed region
        step()
    barrier()
    k1_events.clear()
    launches_before = pipe.k1_launches
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    t_begin.record(stream)
    for _ in range(args.steps):
        step()
    t_end.record(stream)
    barrier()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = t_begin.elapsed_time(t_end)
    k1_ms = sum(a.elapsed_time(b) for a, b in k1_events) / max(args.steps, 1)   # per step: first K1 start to last K1 end (pieces overlap)
    k1_launches = pipe.k1_launches - launches_before
    final_bands = list(pipe.bands)

    # ---- algorithmic flops of this rank's band (untimed counting launch of the same kernel) ------------
    rd.set_scene(sc)
    counters, band_flops = rd.count_rows(width, height, pipe.row0, pipe.row1)

    # N > 1: the assembled stream must be byte-identical to a single-GPU render of the same frame (untimed check)
    stream_ok = None
    if world > 1:
        final = pipe.render(sc)
        if rank == 0:
            import hashlib
            got = hashlib.sha256(final.cpu().numpy().tobytes()).hexdigest()
            rd.use_stream(None)
            want = hashlib.sha256(np.array(rd.render_ansi(sc)).tobytes()).hexdigest()
            rd.use_stream(torch.cuda.current_stream().cuda_stream)
            stream_ok = got == want
        barrier()

    # encoder alone (rank 0's band), for its HBM roofline
    enc_ms = None
    rows = pipe.row1 - pipe.row0
    enc_src = pipe.quant.data_ptr() + (pipe.row0 - pipe.base_row) * width * 4
    if rank == 0 and rows > 0:
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record(stream)
        for _ in range(5):
            rd.encode_rows_quant(enc_src, width, rows, pipe.stream.data_ptr(), abi.HOME_BYTES)
        eb.record(stream)
        torch.cuda.synchronize()
        enc_ms = ea.elapsed_time(eb) / 5

    # ---- end-to-end steps through the host-facing call ----------------------------------------------------
    total_bytes = abi.stream_bytes(width, height)
    scene_bytes = C.sizeof(abi.Scene) + C.sizeof(abi.Sphere) * sc.c.num_spheres + C.sizeof(abi.DirectionalLight) + C.sizeof(abi.PointLight)
    host_pipe = shared = host_out = None
    if world > 1 and pipeline.SharedHostStream.available(total_bytes, rank, world):
        # the stream is wanted in host memory: every rank copies its own bands into one shared page-locked buffer over
        # its own PCIe link (no device-side gather); bands start from the converged ones of the device-resident steps
        shared = pipeline.SharedHostStream(rd, total_bytes, rank, world)
        # (PCIe is ~15x slower than NVLink: more, geometrically shrinking pieces keep the exposed last copy short)
        host_pipe = pipeline.FramePipeline(rd, width, height, rank, world, row_weights=pipe.weights, pieces=(0.4, 0.3, 0.2, 0.1), adapt=True,
                                           host_stream=shared.ptr)     # (fused zero-copy stores over PCIe were measured slower: 11.5 vs 10.6 ms at 8 GPUs)
    elif world > 1 and rank == 0:
        host_out = torch.empty(total_bytes, dtype=torch.uint8).pin_memory()   # no room in /dev/shm: rank 0 copies the gathered stream out

    def e2e_step():
        if world == 1:
            rd.render_ansi(sc)       # trt_render_ansi: H2D scene, K1, K2, D2H stream into pinned memory, sync
        elif host_pipe is not None:
            host_pipe.render(sc)     # set_scene (H2D) + per piece K1, K2, D2H into the shared host stream + closing collective
        else:
            out = pipe.render(sc)    # device-side gather, then one D2H on rank 0
            if rank == 0:
                host_out.copy_(out, non_blocking=True)
            torch.cuda.synchronize()

    if world == 1:
        rd.use_stream(None)          # the plain C-ABI call runs on the library's own stream
    for _ in range(3):
        e2e_step()
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - w0
    host_ok = None
    if world > 1:
        if rank == 0:
            import hashlib
            host_bytes = shared.array.tobytes() if shared is not None else host_out.numpy().tobytes()
            host_ok = hashlib.sha256(host_bytes).hexdigest() == want
        barrier()

    # ---- reduce over ranks ------------------------------------------------------------------------------------
    stats = torch.tensor([ms_total, k1_ms, e2e_s * 1e3, band_flops, k1_launches], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, k1_ms_max, e2e_ms_total = mx[0].item(), mx[1].item(), mx[2].item()
        frame_flops, k1_launches_all, k1_ms_mean = sm[3].item(), int(sm[4].item()), sm[1].item() / world
    else:
        k1_ms_max, e2e_ms_total, frame_flops, k1_launches_all, k1_ms_mean = k1_ms, e2e_s * 1e3, band_flops, k1_launches, k1_ms

    if rank == 0:
        rays_per_step = 10.0 * width * height
        value = rays_per_step * args.steps / (ms_total * 1e-3) / 1e6
        e2e_value = rays_per_step * args.steps / (e2e_ms_total * 1e-3) / 1e6
        peak32, peak64 = peaks["fp32_tflops"], peaks["fp64_tflops"]
        # dominant kernel = K1 on the slowest rank; its algorithmic flops = that launch's band.  With equal-cost
        # accounting across ranks use frame flops / N over the max K1 time (conservative for the roofline).
        achieved = (frame_flops / world) / (k1_ms_max * 1e-3) / 1e12
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                hbm_peak = json.load(f)["hbm_gbs"]
            hbm_src = "MEASURED_PEAKS.json hbm_gbs"
        except (OSError, KeyError, ValueError):
            hbm_peak, hbm_src = 6650.0, "fallback of B200_PROFILING.md"
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(world),
            "frames_per_s": args.steps / (ms_total * 1e-3),
            "roofline": {
                "bound": "alu-fp32", "achieved": achieved, "peak": peak32, "unit": "TFLOP/s", "frac": achieved / peak32,
                "traffic": NCU_DRAM_BYTES["k_render"] if (world == 1 and (width, height) == (WIDTH, HEIGHT)) else None,
                "traffic_unit": "bytes of DRAM traffic per launch (ncu, profiles/r01f_k1_k2_ncu_summary.txt)",
                "kernel": "k_render (K1)", "kernel_ms": k1_ms_max,
                "algorithmic_flops_per_launch": frame_flops / world, "flops_per_primary_ray": frame_flops / rays_per_step,
                "peak_source": "FFMA loop measured in this run by libtrt_b200 (trt_measure_fp32_tflops); MEASURED_PEAKS.json "
                               "has no CUDA-core peak. The kernel executes FP64 (bit-exact parity), whose measured DFMA peak is "
                               f"{peak64:.2f} TFLOP/s",
                "frac_of_fp64_peak": achieved / peak64, "fp64_peak": peak64,
            },
            "roofline_encode": None if enc_ms is None else {
                "bound": "hbm", "achieved": (4.0 * width * rows + abi.row_bytes(width) * rows) / (enc_ms * 1e-3) / 1e9,
                "peak": hbm_peak, "unit": "GB/s", "kernel": "k_encode (K2)", "kernel_ms": enc_ms, "peak_source": hbm_src,
                "frac": (4.0 * width * rows + abi.row_bytes(width) * rows) / (enc_ms * 1e-3) / 1e9 / hbm_peak,
                "traffic": NCU_DRAM_BYTES["k_encode"] if (world == 1 and (width, height) == (WIDTH, HEIGHT)) else None},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(scene_bytes), "d2h_bytes_per_step": int(total_bytes),
                    "ms_per_step": e2e_ms_total / args.steps,
                    "call": "trt_render_ansi(scene,w,h,pinned_out,cap)" if world == 1 else
                            ("FramePipeline(host_stream=shared pinned buffer).render: every rank copies its bands to the host over its own PCIe link"
                             if shared is not None else "FramePipeline.render + D2H of the gathered stream on rank 0 (no room in /dev/shm)")},
            # K1 + K2 per piece on every rank (fused: K1 alone), plus rank 0's trt_stream_frame_device once per step
            "gpu_launches": int((1 if (world > 1 and fused) else 2) * k1_launches_all + args.steps),
            "gather": None if world == 1 else ("fused: K1 stores encoded tiles into rank 0's stream (NVLink peer memory)" if fused else
                                               "pieces: K1, K2, copy-engine push per piece (NVLink peer memory)"),
            "stream_identical_to_single_gpu": stream_ok,
            "host_stream_identical_to_single_gpu": host_ok,
            "bands": None if world == 1 else {"rows": final_bands, "k1_ms_max_rank": k1_ms_max, "k1_ms_mean_rank": k1_ms_mean,
                                              "how": "1/8-resolution cost pre-pass, then feedback from the ranks' measured K1 times of the previous steps"},
            "clocks": clocks,
            "work_counters_rank0": {"trace_calls": counters[9], "sphere_tests": counters[0], "sky_lookups": counters[8],
                                    "bounce_iters": counters[12], "lighting_calls": counters[11]},
        }
        if not args.no_cpu_baseline:
            kind, times = time_reference_cpu(CPU_SAMPLE_W, CPU_SAMPLE_H, 12)
            best = min(times[1:])
            mean = sum(times[1:]) / len(times[1:])
            line["cpu_baseline"] = {
                "value": 10.0 * CPU_SAMPLE_W * CPU_SAMPLE_H / mean / 1e6, "unit": "Mrays/s", "cores": 1, "kind": kind,
                "best_value": 10.0 * CPU_SAMPLE_W * CPU_SAMPLE_H / best / 1e6, "host_cores_available": os.cpu_count(),
                "sample": f"11 timed frames of the same scene, pose and skybox at {CPU_SAMPLE_W}x{CPU_SAMPLE_H} (1/256 of the pixels), "
                          f"single thread as the reference is written, gcc -O3 -ffp-contract=off"}
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        if shared is not None:
            shared.close()
        if rank != 0:
            pipe.close()             # importers release rank 0's buffer before rank 0 frees it
        dist.barrier()
        pipe.close()
        dist.destroy_process_group()
    rd.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
