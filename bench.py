#!/usr/bin/env python
"""bench.py — headline benchmark of the render path (BASELINE.json: Mrays/s on the default demo scene
at 7680x4320, reference bounce depth, milky_way skybox, row-band sharded over 1/2/4/8 GPUs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config demo8k|demo4k|stress|orbit]
                    [--fused -1|0|1|2] [--pieces 0.7,0.3]        (gather form and piece fractions at N > 1, see --help)
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

# BASELINE.json `configs`: [2] is the headline (default); the others are selected with --config and are measured with the
# same harness (profiles/r02_config*.json); [0] is the reference's own CPU case and has no GPU line.
CONFIGS = {
    "demo8k": {"label": "BASELINE config 2", "kind": "demo", "width": 7680, "height": 4320, "skybox": "milky_way", "t": 3.7,
               "cpu_rows": list(range(128, 4320, 256))},
    "demo4k": {"label": "BASELINE config 1", "kind": "demo", "width": 3840, "height": 2160, "skybox": "uv_checker", "t": 3.7,
               "cpu_rows": list(range(64, 2160, 128))},
    "stress": {"label": "BASELINE config 3", "kind": "stress", "spheres": 1024, "width": 3840, "height": 2160, "skybox": "uv_checker", "t": 3.7,
               "cpu_rows": [270, 810, 1350, 1890]},
    "orbit": {"label": "BASELINE config 4", "kind": "demo", "width": 1920, "height": 1080, "skybox": "milky_way", "frames": 360,
              "cpu_frames": [0, 120, 240], "cpu_rows": list(range(32, 1080, 64))},
}
TRAFFIC_STAMP = os.path.join(ROOT, "profiles", "k1_k2_traffic_stamp.json")


def workload_config(cfg, n_gpus):
    w, h = cfg["width"], cfg["height"]
    if cfg["kind"] == "stress":
        what = f"synthetic stress scene: {cfg['spheres']} random spheres with mixed materials (splitmix64 seed 0x5EED1024, SURVEY 8d) + checker ground + 1 directional + 1 point light"
    else:
        what = "default demo scene (6 spheres + checker ground + 1 directional + 1 point light)"
    sky = cfg["skybox"] + (" (synthetic 1024^2 x 6 stand-in: the reference's milky_way assets are not in its checkout)" if cfg["skybox"] == "milky_way" else " (the reference's own asset)")
    if "frames" in cfg:
        pose = f"{cfg['frames']}-frame camera orbit, t_k = k*20/{cfg['frames']} s (one yaw turn at the reference's 0.05 rev/s)"
        shard = (f"frame k on rank k mod {n_gpus}; every rank copies its frames from the device into one shared page-locked ring over its own "
                 f"PCIe link, rank 0 writes them out strictly in order (no gather, no data-path collective)") if n_gpus > 1 else "single GPU"
    else:
        pose = f"orbit pose t={cfg['t']}s"
        shard = (f"cost-weighted contiguous row-bands x{n_gpus} (1/8-resolution cost pre-pass, then feedback from the ranks' measured K1 "
                 f"times); every rank's encoded bytes go into rank 0's stream over NVLink peer memory (see 'gather'); a step ends on the device "
                 f"(per-rank flag words behind rank 0's stream, no host synchronisation); a small NCCL all-gather of the K1 times runs only on "
                 f"the feedback steps (the first five, then every 32nd)") if n_gpus > 1 else "single GPU"
    return {
        "workload": f"{cfg['label']}: {what}, {w}x{h} cells, 10 samples/pixel, bounce limit 10, skybox {sky}, {pose}",
        "width": w, "height": h, "samples_per_pixel": 10, "bounce_limit": 10, "skybox": cfg["skybox"],
        "sharding": shard,
        "l2": f"no explicit flush: each step writes {4 * w * h / 1e6:.0f} MB of cells + {(25 * w + 1) * h / 1e6:.0f} MB of stream"
              + (" (> 126 MB L2)" if 29 * w * h > 126e6 else " per frame, a different camera pose every frame") +
              "; inputs (scene 1 KB, skybox 25 MB) are meant to stay cache resident, the kernel is bound by instruction issue, not memory",
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

def _stress_spheres(n):
    """the stress scene's spheres as trt_stress_scene (csrc/trt_host.c) generates them — restated here so that the reference
    arm builds its input without mapping the product library (tests/test_host.py checks the two agree bit for bit)"""
    from terminalraytracer_b200 import abi
    mask = (1 << 64) - 1
    state = [0x5EED1024]

    def unit():
        state[0] = (state[0] + 0x9E3779B97F4A7C15) & mask
        z = state[0]
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & mask
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & mask
        z ^= z >> 31
        return float(z >> 11) * (1.0 / 9007199254740992.0)

    def rng(lo, hi):
        return lo + unit() * (hi - lo)

    reflect = (0.0, 0.2, 0.8, 1.0)
    out = (abi.Sphere * n)()
    made = 0
    while made < n:
        cx, cy, cz = rng(-8.0, 8.0), rng(-1.5, 6.0), rng(-8.0, 8.0)
        radius = rng(0.1, 0.5)
        r, g, b = rng(0.0, 1.0), rng(0.0, 1.0), rng(0.0, 1.0)
        d = (cx * cx + cy * cy + cz * cz) ** 0.5
        if 1.99 - radius - 0.05 < d < 1.99 + radius + 0.05:
            continue
        out[made] = abi.Sphere(abi.Vector(cx, cy, cz), radius, abi.Material(abi.Vector(r, g, b), reflect[made & 3], 100.0))
        made += 1
    return out


class ReferenceCPU:
    """The reference's own project_scene on ONE host thread (it is single-threaded as written), on rows sampled from the REAL
    frame of the configured workload: oracle/_ref/libtrt_ref_rows.so is the unmodified TU with the row loop of TRT.c:973
    bounded (oracle/Makefile); the scene comes from the reference's init_camera, the literals of its main() and its camera
    recipe (oracle/ref_harness.c: ref_demo_scene, ref_orbit_camera).  The product library is never loaded here.  Falls back to
    the oracle port (kind "port") when the reference build did not travel to the box."""

    def __init__(self, cfg):
        from terminalraytracer_b200 import abi, scene as S
        self.cfg, self.abi = cfg, abi
        root = os.path.join(ROOT, "oracle")
        ref_rows, ref_so = os.path.join(root, "_ref", "libtrt_ref_rows.so"), os.path.join(root, "_ref", "libtrt_ref.so")
        if os.path.exists(ref_rows):
            self.lib, self.kind = C.CDLL(ref_rows), "reference"
            self.render = self.lib.ref_project_rows
            self.render.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.Screen), C.c_int, C.c_int]
        else:
            port = os.path.join(root, "_build", "libtrt_oracle.so")
            if not os.path.exists(port):
                subprocess.check_call(["make", "-s", "-C", root, "oracle"])
            self.lib, self.kind = C.CDLL(port), "port"
            self.lib.orc_render_rows.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.Screen), C.c_int, C.c_int, C.c_void_p]
            self.render = lambda sc, scr, r0, r1: self.lib.orc_render_rows(sc, scr, r0, r1, None)
        w, h = cfg["width"], cfg["height"]
        self.sky = S.get_skybox(cfg["skybox"])              # numpy + ctypes only
        self.scene = abi.Scene()
        self.scene.skybox = self.sky.c
        self.dl, self.pl = abi.DirectionalLight(), abi.PointLight()
        self.spheres = (abi.Sphere * 6)()
        builder = C.CDLL(ref_so) if os.path.exists(ref_so) else None
        if builder is not None:
            builder.ref_demo_scene.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.Sphere), C.POINTER(abi.DirectionalLight),
                                               C.POINTER(abi.PointLight), C.c_int, C.c_int]
            builder.ref_orbit_camera.argtypes = [C.POINTER(abi.Camera), C.c_double]
            builder.ref_demo_scene(C.byref(self.scene), self.spheres, C.byref(self.dl), C.byref(self.pl), w, h)
            self.pose = lambda t: builder.ref_orbit_camera(C.byref(self.scene.camera), float(t))
            self.scene_by = "the reference's init_camera / main() literals / camera recipe (oracle/_ref/libtrt_ref.so)"
        else:
            # no reference build on this box: the product's host C helpers (bit-identical, tests/test_host.py) build the scene
            sd = S.SceneData(w, h, self.sky)
            self.keep = sd
            self.scene, self.spheres = sd.c, sd.spheres
            self.pose = lambda t: sd.set_time(t)
            self.scene_by = "libtrt_b200's host C helpers (no reference build on this box)"
        if cfg["kind"] == "stress":
            self.spheres = _stress_spheres(cfg["spheres"])
            self.scene.spheres = C.cast(self.spheres, C.POINTER(abi.Sphere))
            self.scene.num_spheres = cfg["spheres"]
        self.rows = [r for r in cfg["cpu_rows"] if r < h]
        self.frames = [sharding_time(cfg, k) for k in cfg.get("cpu_frames", [None])]
        import numpy as np
        self.px = np.zeros((len(self.rows), w, 3))

    def rays_per_pass(self):
        return 10.0 * self.cfg["width"] * len(self.rows) * len(self.frames)

    def one_pass(self):
        """every sampled row of every sampled pose once; returns seconds"""
        w, h = self.cfg["width"], self.cfg["height"]
        t0 = time.perf_counter()
        for t in self.frames:
            self.pose(t)
            for i, r in enumerate(self.rows):
                base = self.px[i:].ctypes.data - r * w * 24     # a Screen whose row r lands at px[i]
                scr = self.abi.Screen(C.cast(C.c_void_p(base), C.POINTER(self.abi.Vector)), w, h)
                self.render(C.byref(self.scene), C.byref(scr), r, r + 1)
        return time.perf_counter() - t0

    def sample_text(self, passes):
        cfg = self.cfg
        where = (f"{len(self.rows)} rows (every {self.rows[1] - self.rows[0]}th, first {self.rows[0]}) of the real "
                 f"{cfg['width']}x{cfg['height']} frame" if len(self.rows) > 1 else f"row {self.rows[0]}")
        if "frames" in cfg:
            where += f" at frames {cfg['cpu_frames']} of the {cfg['frames']}-frame orbit"
        return (f"{passes} timed passes over {where} = {self.rays_per_pass() / 1e6:.3f} M primary rays per pass; project_scene of the "
                f"{'unmodified reference TU (row loop bounded, oracle/Makefile)' if self.kind == 'reference' else 'oracle port'}, "
                f"gcc -O3 -ffp-contract=off, 1 thread as the reference is written; scene built by {self.scene_by}")


def sharding_time(cfg, k):
    return cfg["t"] if k is None else k * (20.0 / cfg["frames"])


def run_reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cpu = ReferenceCPU(cfg)
    times = [cpu.one_pass() for _ in range(args.warmup + args.steps)][args.warmup:]
    total = sum(times)
    value = cpu.rays_per_pass() * len(times) / total / 1e6
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(cfg, args.gpus),
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": 1, "kind": cpu.kind, "sample": cpu.sample_text(len(times)),
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if "frames" in cfg:
        line["frames_per_s"] = value * 1e6 / (10.0 * cfg["width"] * cfg["height"])
    print(json.dumps(line))
    return 0


def cpu_baseline_leg(cfg, budget_s=20.0):
    """cpu_baseline of the product arm: the same sampler, about budget_s seconds of single-thread CPU work"""
    cpu = ReferenceCPU(cfg)
    first = cpu.one_pass()                       # warm-up pass, also sizes the loop
    passes = max(2, min(12, int(budget_s / max(first, 1e-3))))
    times = [cpu.one_pass() for _ in range(passes)]
    mean, best = sum(times) / len(times), min(times)
    out = {"value": cpu.rays_per_pass() / mean / 1e6, "unit": "Mrays/s", "cores": 1, "kind": cpu.kind,
           "best_value": cpu.rays_per_pass() / best / 1e6, "host_cores_available": os.cpu_count(), "sample": cpu.sample_text(passes)}
    if "frames" in cfg:
        out["frames_per_s"] = out["value"] * 1e6 / (10.0 * cfg["width"] * cfg["height"])
    return out


def host_pieces(world):
    """Piece schedule of the host-stream path (e2e, N > 1): a rank's band leaves for the host piece by piece, each copy behind
    the next piece's K1.  With N >= 2 the copies are the floor (829 MB into one host buffer: ~10.5 ms at the ~80 GB/s the host
    ingests, scripts/pcie_bw.py; K1 per rank is 23 ms / N), so the copy engine has to start early and never wait: many pieces,
    small ones first (the first copy starts after ~4 % of the band) and last (the exposed tail).  Measured at N = 2 (one box,
    back to back, profiles/r02_e2e_pieces_n2.txt): 4 pieces 0.4/0.3/0.2/0.1 (round 1) 17.5 ms, 4 even-ish 16.7, 5 pieces 16.2,
    6 pieces 15.6, 8 pieces 14.9."""
    if os.environ.get("TRT_HOST_PIECES"):            # experiments: "0.1,0.3,0.3,0.2,0.1"
        return tuple(float(x) for x in os.environ["TRT_HOST_PIECES"].split(","))
    return (0.04, 0.08, 0.12, 0.14, 0.14, 0.14, 0.12, 0.1, 0.08, 0.04)


def traffic_stamp():
    """DRAM bytes per launch of K1 / K2 from the last `ncu --set full` capture of the bench workload, stamped with the hash of the
    kernel sources it was taken at (scripts/stamp_traffic.py); `current` says whether the kernels are still those."""
    import hashlib
    try:
        with open(TRAFFIC_STAMP) as f:
            st = json.load(f)
    except (OSError, ValueError):
        return None
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "terminalraytracer_b200", "csrc")
    for name in sorted(os.listdir(csrc)):
        if name.endswith((".cu", ".cuh", ".h")):
            with open(os.path.join(csrc, name), "rb") as f:
                h.update(f.read())
    st["current"] = st.get("kernel_sources_sha256") == h.hexdigest()
    return st


# --------------------------------------------------------------------------------------------------------
# BASELINE config 4: the camera orbit, frame-sharded, streamed in order

def run_orbit(args, cfg):
    """--config orbit.  A step is one frame of the orbit; --steps K renders frames 0..K-1 of the 360-frame path (default: all 360),
    frame k on rank k mod N ("scaling": strong — the animation is fixed, the ranks share it).
      value  frames/s with every rank's frames rendered and encoded on the device (K1 + K2 per frame, camera re-posed and the
             scene constants uploaded every frame), device-timed per rank, max over ranks.
      e2e    frames/s through OrbitPipeline.stream: trt_render_orbit_to on every rank, device -> shared page-locked ring over
             the rank's own PCIe link, rank 0 writes every frame to /dev/null strictly in order (one write per frame, TRT.c:1171);
             wall clock between barriers."""
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import hashlib
    import numpy as np
    import torch
    import torch.distributed as dist
    from terminalraytracer_b200 import abi, pipeline, renderer as R, scene as S, sharding

    width, height = cfg["width"], cfg["height"]
    rank, world, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
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

    n_frames = cfg["frames"] if args.steps is None else args.steps
    times = [k * (20.0 / cfg["frames"]) for k in range(n_frames)]
    mine = sharding.frames_for_rank(n_frames, rank, world)
    rd = R.Renderer(local_rank)
    sky = S.get_skybox(cfg["skybox"])
    rd.upload_skybox(sky)
    peaks = rd.measure_peaks() if rank == 0 else None
    sc = S.SceneData(width, height, sky)
    total_bytes = abi.stream_bytes(width, height)
    stream = torch.cuda.current_stream()
    rd.use_stream(stream.cuda_stream)
    quant = torch.empty(width * height * 4, dtype=torch.uint8, device="cuda")
    dev_stream = torch.empty(total_bytes + 16, dtype=torch.uint8, device="cuda")

    k1_pairs = []

    def device_frame(k, timed):
        sc.set_time(times[k])
        rd.set_scene_async(sc)                               # H2D of the re-posed scene, no host wait
        a = b = None
        if timed:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
        rd.render_rows_quant(width, height, 0, height, quant.data_ptr())
        if timed:
            b.record(stream)
            k1_pairs.append((a, b))
        rd.stream_frame(dev_stream.data_ptr(), width, height)
        rd.encode_rows_quant(quant.data_ptr(), width, height, dev_stream.data_ptr(), abi.HOME_BYTES)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for k in mine[:max(args.warmup, 3)]:
        device_frame(k, False)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark_begin()
    t0.record(stream)
    for k in mine:
        device_frame(k, True)
    t1.record(stream)
    barrier()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = t0.elapsed_time(t1)
    k1_ms = sum(a.elapsed_time(b) for a, b in k1_pairs)

    # algorithmic flops: three poses of the path, counted by the kernel itself (untimed)
    flops = []
    if rank == 0:
        for k in sorted(set(int(f * n_frames / cfg["frames"]) for f in cfg["cpu_frames"] if int(f * n_frames / cfg["frames"]) < n_frames)):
            sc.set_time(times[k])
            rd.set_scene(sc)
            flops.append(rd.count_rows(width, height, 0, height)[1])
    rd.use_stream(None)

    # ---- end to end: ordered stream to /dev/null ----------------------------------------------------------------
    orbit = pipeline.OrbitPipeline(rd, width, height, rank, world)
    devnull = os.open(os.devnull, os.O_WRONLY)

    def write(k, view):
        return os.write(devnull, view) != len(view)

    warm = [k * (20.0 / cfg["frames"]) for k in range(min(n_frames, 4 * world))]
    orbit.stream(S.SceneData(width, height, sky), warm, write)
    barrier()
    w0 = time.perf_counter()
    done, written = orbit.stream(S.SceneData(width, height, sky), times, write)
    barrier()
    e2e_s = time.perf_counter() - w0
    os.close(devnull)

    # untimed check: the first frames of the ordered stream are byte-identical to single-GPU renders of the same poses
    check_n = min(n_frames, max(2 * world, 4))
    got = []
    orbit.stream(S.SceneData(width, height, sky), times[:check_n], lambda k, view: got.append(hashlib.sha256(view).hexdigest()) and False)
    ordered_ok = None
    if rank == 0:
        want = []
        one = S.SceneData(width, height, sky)
        for k in range(check_n):
            one.set_time(times[k])
            want.append(hashlib.sha256(np.array(rd.render_ansi(one)).tobytes()).hexdigest())
        ordered_ok = got == want
    barrier()

    stats = torch.tensor([dev_ms, k1_ms, e2e_s * 1e3, float(done)], dtype=torch.float64, device="cuda")
    mx, sm = stats.clone(), stats.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    if rank == 0:
        dev_ms_max, k1_ms_max, e2e_ms = mx[0].item(), mx[1].item(), mx[2].item()
        fps, e2e_fps = n_frames / (dev_ms_max * 1e-3), n_frames / (e2e_ms * 1e-3)
        rays = 10.0 * width * height
        flops_per_frame = sum(flops) / max(len(flops), 1)
        frames_slowest = len(sharding.frames_for_rank(n_frames, 0, world))
        achieved = flops_per_frame * frames_slowest / (k1_ms_max * 1e-3) / 1e12
        peak32, peak64 = peaks["fp32_tflops"], peaks["fp64_tflops"]
        line = {
            "metric": "frames/s", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": n_frames, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms_max / max(frames_slowest, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(cfg, world),
            "mrays_per_s": fps * rays / 1e6,
            "frames_rendered_all_ranks": int(sm[3].item()), "frames_written_in_order": int(written),
            "roofline": {
                "bound": "alu-fp32 (issue slots of the CUDA cores)", "achieved": achieved, "peak": peak32, "unit": "TFLOP/s", "frac": achieved / peak32,
                "achieved_is": "ALGORITHMIC flops (SURVEY 8d counter model, mean of three poses of the path) x frames of a rank / that rank's summed K1 time",
                "traffic": None, "kernel": "k_render (K1) incl. k_tile_certs", "kernel_ms": k1_ms_max / max(frames_slowest, 1),
                "flops_per_primary_ray": flops_per_frame / rays, "frac_of_fp64_peak": achieved / peak64, "fp64_peak": peak64,
                "peak_source": "FFMA / DFMA loops measured in this run by libtrt_b200"},
            "e2e": {"value": e2e_fps, "unit": "frames/s", "mrays_per_s": e2e_fps * rays / 1e6, "ms_per_frame": e2e_ms / n_frames,
                    "h2d_bytes_per_step": int(C.sizeof(abi.Scene) + C.sizeof(abi.Sphere) * 6 + C.sizeof(abi.DirectionalLight) + C.sizeof(abi.PointLight)),
                    "d2h_bytes_per_step": int(total_bytes), "host_ingest_gb_per_s": e2e_fps * total_bytes / 1e9,
                    "call": "OrbitPipeline.stream: trt_render_orbit_to(first=rank, stride=N) per rank into the shared page-locked ring, rank 0 writes "
                            "every frame to /dev/null in order while later frames render"},
            # per frame: k_tile_certs, k_render, k_stream_frame, k_encode — in the device-timed region
            "gpu_launches": 4 * n_frames,
            "ordered_stream_identical_to_single_gpu": ordered_ok, "frames_checked": check_n,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_leg(cfg)
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    barrier()
    orbit.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    rd.close()
    return 0


# --------------------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default: 20 frames (frame configs) / the whole 360-frame path (orbit)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="demo8k", choices=sorted(CONFIGS),
                    help="BASELINE.json configs: demo8k = [2] (default, the headline), demo4k = [1], stress = [3], orbit = [4]")
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=None, help="steps of the end-to-end leg (default: --steps)")
    ap.add_argument("--fused", type=int, default=int(os.environ.get("TRT_BENCH_FUSED", "-1")),
                    help="N > 1, device-resident gather: 1 = K1 stores the encoded tiles straight into rank 0's stream over NVLink (no K2, "
                         "no copies), 0 = K1, K2 and copy-engine pushes piece by piece, 2 = K1 then ONE K2 per band that stores into rank 0's stream "
                         "through the peer mapping, -1 = by GPU count (measured, profiles/r02f_*: pieces and direct within 1 %% at 2 and 4 GPUs, "
                         "12.18 / 12.12 and 6.28 / 6.35 ms; the fused kernel wins at 8, 3.45 vs 3.63 ms)")
    ap.add_argument("--pieces", default=os.environ.get("TRT_BENCH_PIECES", "0.7,0.3"),
                    help="N > 1, pieces and direct gather: the fractions of a rank's band rendered, encoded and sent on their way one after the other")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.width and args.height and (args.width, args.height) != (cfg["width"], cfg["height"]):
        cfg.update(width=args.width, height=args.height, label=cfg["label"] + " (resized on the command line)",
                   cpu_rows=[r for r in range(0, args.height, max(1, args.height // 16))])
    if args.impl == "reference":
        if args.steps is None:
            args.steps = 20 if "frames" not in cfg else 5
        return run_reference_arm(args, cfg)
    if "frames" in cfg and args.impl != "reference":
        return run_orbit(args, cfg)
    if args.steps is None:
        args.steps = 20

    # stdout carries exactly one JSON line: whatever libraries print there (NCCL's version banner) goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import torch.distributed as dist
    from terminalraytracer_b200 import abi, pipeline, renderer as R, scene as S

    width, height = cfg["width"], cfg["height"]
    headline = args.config == "demo8k" and (width, height) == (CONFIGS["demo8k"]["width"], CONFIGS["demo8k"]["height"])
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
    sky = S.get_skybox(cfg["skybox"])
    rd.upload_skybox(sky)
    sc = S.SceneData(width, height, sky, kind=cfg["kind"], num_spheres=cfg.get("spheres", 1024)).set_time(cfg["t"])
    # cost-weighted row bands (sky rows are ~5x cheaper than sphere/ground rows): every rank runs the same
    # deterministic 1/8-resolution pre-pass and derives the same bands; untimed, once per scene
    weights = rd.estimate_row_costs(sc) if world > 1 else None
    fused = (world >= 8) if args.fused < 0 else args.fused == 1
    direct = args.fused == 2
    # N > 1: every rank pushes its encoded pieces into rank 0's stream over NVLink peer memory while its next piece renders
    # (or, fused: K1 itself stores every finished tile's bytes there);
    # the collective that ends a step carries the ranks' K1 times and the next step's bands follow from them (adapt)
    pipe = pipeline.FramePipeline(rd, width, height, rank, world, row_weights=weights, peer=world > 1, pieces=tuple(float(x) for x in args.pieces.split(",")), adapt=world > 1, fused=fused, direct=direct)
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
    t_end.record(stream)      # rank 0's stream ends with the device-side wait for every rank's last step
    pipe.finish()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = t_begin.elapsed_time(t_end)
    k1_ms = sum(a.elapsed_time(b) for a, b in k1_events) / max(args.steps, 1)   # per step: first K1 start to last K1 end (pieces overlap)
    k1_launches = pipe.k1_launches - launches_before
    final_bands = list(pipe.bands)

    # ---- algorithmic flops of this rank's band (untimed counting launch of the same kernel) ------------
    rd.set_scene(sc)
    flops_note = "work counters of the whole band (counting flavour of K1, untimed)"
    if cfg["kind"] == "stress" and pipe.row1 - pipe.row0 > 8:
        # the counting flavour answers every query the reference's way as well (1024 exact sphere tests per query): count 8 evenly
        # spaced rows of the band and scale — the as-written model is only a yardstick here, the kernel skips most of that work
        rows = sorted(set(pipe.row0 + (i * (pipe.row1 - pipe.row0 - 1)) // 7 for i in range(8)))
        counters, band_flops = [0] * abi.NUM_COUNTERS, 0.0
        for r in rows:
            c, f = rd.count_rows(width, height, r, r + 1)
            counters = [a + b for a, b in zip(counters, c)]
            band_flops += f
        scale = (pipe.row1 - pipe.row0) / len(rows)
        counters, band_flops = [int(c * scale) for c in counters], band_flops * scale
        flops_note = f"work counters of {len(rows)} evenly spaced rows of the band, scaled by rows (the counting flavour tests all 1024 spheres per query)"
    else:
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
        host_pipe = pipeline.FramePipeline(rd, width, height, rank, world, row_weights=pipe.weights, pieces=host_pieces(world), adapt=True,
                                           host_stream=shared.ptr, host_sync=shared)     # (fused zero-copy stores over PCIe were measured slower: 11.5 vs 10.6 ms at 8 GPUs)
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
    e2e_steps = args.e2e_steps or args.steps
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - w0) * (args.steps / e2e_steps)      # per-step time x steps, as the reduction below expects
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
        stamp = traffic_stamp() if (world == 1 and headline) else None
        k1_traffic = stamp["k_render_dram_bytes_per_launch"] if stamp else None
        executed = None
        if stamp and stamp.get("k_render_executed"):
            ex = stamp["k_render_executed"]
            executed = {"fp64_flop_per_launch": ex["fp64_flop"], "fp32_flop_per_launch": ex["fp32_flop"],
                        "fp64_tflops": ex["fp64_flop"] / (k1_ms_max * 1e-3) / 1e12, "fp32_tflops": ex["fp32_flop"] / (k1_ms_max * 1e-3) / 1e12,
                        "frac_of_fp64_peak": ex["fp64_flop"] / (k1_ms_max * 1e-3) / 1e12 / peak64,
                        "warp_instructions": ex.get("warp_instructions"), "issue_active_pct": ex.get("issue_active_pct"),
                        "what": "arithmetic the kernel really executes (ncu sass counters of the stamped capture; packed FFMA2/FADD2/FMUL2 "
                                "certificate arithmetic is not in the scalar FP32 counters): the certificates rule out most of the "
                                "reference's as-written sphere tests, so this is well below the algorithmic figure by design"}
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(cfg, world),
            "frames_per_s": args.steps / (ms_total * 1e-3),
            "roofline": {
                "bound": "alu-fp32 (issue slots of the CUDA cores; no dense contraction, ~400 flop/byte)", "achieved": achieved, "peak": peak32,
                "unit": "TFLOP/s", "frac": achieved / peak32,
                "achieved_is": "ALGORITHMIC flops (the reference's as-written work, SURVEY 8d counter model) per launch / kernel time — "
                               "not executed flops, see 'executed'",
                "traffic": k1_traffic,
                "traffic_source": None if not stamp else {k: stamp[k] for k in ("profile", "commit", "kernel_sources_sha256", "current") if k in stamp},
                "traffic_unit": "dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full)",
                "kernel": "k_render (K1) incl. its tile-certificate prepass k_tile_certs", "kernel_ms": k1_ms_max,
                "algorithmic_flops_per_launch": frame_flops / world, "flops_per_primary_ray": frame_flops / rays_per_step,
                "flops_counted_by": flops_note,
                "peak_source": "FFMA loop measured in this run by libtrt_b200 (trt_measure_fp32_tflops); MEASURED_PEAKS.json "
                               "has no CUDA-core peak. The kernel executes FP64 (bit-exact parity), whose measured DFMA peak is "
                               f"{peak64:.2f} TFLOP/s",
                "frac_of_fp64_peak": achieved / peak64, "fp64_peak": peak64,
                "executed": executed,
            },
            "roofline_encode": None if enc_ms is None else {
                "bound": "hbm", "achieved": (4.0 * width * rows + abi.row_bytes(width) * rows) / (enc_ms * 1e-3) / 1e9,
                "peak": hbm_peak, "unit": "GB/s", "kernel": "k_encode (K2)", "kernel_ms": enc_ms, "peak_source": hbm_src,
                "frac": (4.0 * width * rows + abi.row_bytes(width) * rows) / (enc_ms * 1e-3) / 1e9 / hbm_peak,
                "traffic": stamp["k_encode_dram_bytes_per_launch"] if stamp else None},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(scene_bytes), "d2h_bytes_per_step": int(total_bytes),
                    "ms_per_step": e2e_ms_total / args.steps,
                    "call": "trt_render_ansi(scene,w,h,pinned_out,cap)" if world == 1 else
                            ("FramePipeline(host_stream=shared pinned buffer).render: every rank copies its bands to the host over its own PCIe link"
                             if shared is not None else "FramePipeline.render + D2H of the gathered stream on rank 0 (no room in /dev/shm)")},
            # per piece on every rank: k_tile_certs (small scenes) + K1 (+ K2 unless fused), plus rank 0's trt_stream_frame_device once per step
            # N > 1 adds the one-thread flag kernels of the device-side step completion: every rank's k_signal, rank 0's k_wait_flags +
            # k_signal ("consumed"), and the other ranks' k_wait_flags in front of their first write of the next frame
            "gpu_launches": int(((1 if (world > 1 and fused) else 2) + (1 if cfg["kind"] == "demo" else 0)) * k1_launches_all + args.steps
                                + ((2 * world + 1) * args.steps if world > 1 else 0)),
            "gather": None if world == 1 else ("fused: K1 stores encoded tiles into rank 0's stream (NVLink peer memory)" if fused else
                                               ("direct: K1, then K2 stores the piece's bytes into rank 0's stream (NVLink peer memory), pieces " + args.pieces if direct else
                                                "pieces: K1, K2, copy-engine push per piece (NVLink peer memory), pieces " + args.pieces)),
            "stream_identical_to_single_gpu": stream_ok,
            "host_stream_identical_to_single_gpu": host_ok,
            "bands": None if world == 1 else {"rows": final_bands, "k1_ms_max_rank": k1_ms_max, "k1_ms_mean_rank": k1_ms_mean,
                                              "how": "1/8-resolution cost pre-pass, then feedback from the ranks' measured K1 times of the previous steps"},
            "clocks": clocks,
            "work_counters_rank0": {"trace_calls": counters[9], "sphere_tests": counters[0], "sky_lookups": counters[8],
                                    "bounce_iters": counters[12], "lighting_calls": counters[11]},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_leg(cfg)
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
