"""The N>1 path on CPU: world_size-2 and -3 gloo process groups run the product's gather logic
(terminalraytracer_b200/dist.py + sharding.py) on bands/frames whose bytes come from the oracle — the
checker supplies the payload here; the code under test is the partition and the exchange."""
import ctypes as C
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from terminalraytracer_b200 import abi, dist as tdist, scene as S, sharding
from tests import _util as U


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _band_worker(rank, world, port, w, h, weights, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = U.load_oracle()
        sc = S.SceneData(w, h, S.synthetic_cubemap("uv_gradient", 64)).set_time(3.7)
        bands = sharding.row_bands(h, world, weights)
        r0, r1 = bands[rank]
        # each rank renders and encodes ONLY its band
        px = np.zeros((h, w, 3))
        scr = U.screen_for(px)
        orc.orc_render_rows(C.byref(sc.c), C.byref(scr), r0, r1, None)
        band = np.zeros(max((r1 - r0) * abi.row_bytes(w), 1), dtype=np.uint8)
        n = orc.orc_encode_rows(C.byref(scr), r0, r1, U.VP(band.ctypes.data))
        assert n == (r1 - r0) * abi.row_bytes(w)
        band_t = torch.from_numpy(band[:n])
        stream = None
        if rank == 0:
            stream = torch.zeros(abi.stream_bytes(w, h), dtype=torch.uint8)
            stream[:6] = torch.tensor(list(b"\033[0;0H"), dtype=torch.uint8)
            b0, b1 = sharding.band_byte_range(w, bands[0])
            stream[b0:b1] = band_t
        got = tdist.gather_bands(stream, band_t, w, bands, rank, world)
        if rank == 0:
            np.save(out_path, got.numpy())
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


def _run_bands(world, w, h, weights, tmp_path):
    out = str(tmp_path / f"stream_{world}.npy")
    mp.spawn(_band_worker, args=(world, _free_port(), w, h, weights, out), nprocs=world, join=True)
    orc = U.load_oracle()
    sc = S.SceneData(w, h, S.synthetic_cubemap("uv_gradient", 64)).set_time(3.7)
    want = U.oracle_stream(orc, U.cpu_render(orc, "orc_project_scene", sc))
    assert np.array_equal(np.load(out), want)


def test_row_band_gather_world2(tmp_path):
    _run_bands(2, 33, 17, None, tmp_path)


def test_row_band_gather_world3_weighted_with_empty_band(tmp_path):
    # weights push all rows onto two ranks' bands; the third may be tiny or empty
    h = 9
    weights = [10.0] * 2 + [0.1] * 7
    _run_bands(3, 20, h, weights, tmp_path)


def _frame_worker(rank, world, port, nframes, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = U.load_oracle()
        w, h = 16, 9
        sc = S.SceneData(w, h, S.synthetic_cubemap("colors", 16))
        times = sharding.orbit_times(nframes)
        ids = sharding.frames_for_rank(nframes, rank, world)
        mine = []
        for k in ids:
            sc.set_time(times[k])
            mine.append(torch.from_numpy(U.oracle_stream(orc, U.cpu_render(orc, "orc_project_scene", sc))))
        got = tdist.gather_frames(mine, ids, nframes, abi.stream_bytes(w, h), rank, world, torch.device("cpu"))
        if rank == 0:
            np.save(out_path, np.stack([g.numpy() for g in got]))
    finally:
        dist.destroy_process_group()


def test_frame_sharded_gather_world2(tmp_path):
    out = str(tmp_path / "frames.npy")
    nframes = 5
    mp.spawn(_frame_worker, args=(2, _free_port(), nframes, out), nprocs=2, join=True)
    orc = U.load_oracle()
    sc = S.SceneData(16, 9, S.synthetic_cubemap("colors", 16))
    got = np.load(out)
    for k, t in enumerate(sharding.orbit_times(nframes)):
        sc.set_time(t)
        want = U.oracle_stream(orc, U.cpu_render(orc, "orc_project_scene", sc))
        assert np.array_equal(got[k], want), k


# ---- the ordered sink of the frame-sharded animation (pipeline.OrderedFrameRing): host logic, no GPU ----------------

def _ring_worker(rank, world, port, nframes, slots, out_path):
    import time
    from terminalraytracer_b200 import pipeline
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = U.load_oracle()
        w, h = 16, 9
        nbytes = abi.stream_bytes(w, h)
        ring = pipeline.OrderedFrameRing(nbytes, nframes, slots, rank, world)
        order, frames = [], []
        consumer = None
        if rank == 0:
            import threading

            def write(k, view):
                order.append(k)
                frames.append(np.frombuffer(view, dtype=np.uint8).copy())
                time.sleep(0.002)                   # a slow terminal: producers must wait for free slots
                return False

            consumer = threading.Thread(target=lambda: ring.consume(write))
            consumer.start()
        sc = S.SceneData(w, h, S.synthetic_cubemap("colors", 16))
        times = sharding.orbit_times(nframes)
        for k in sharding.frames_for_rank(nframes, rank, world):
            if rank == world - 1:
                time.sleep(0.004)                   # the last rank is late: the consumer must wait for ITS frames, in order
            addr = ring.acquire(k)
            assert addr == ring.slot_address(k)
            sc.set_time(times[k])
            data = U.oracle_stream(orc, U.cpu_render(orc, "orc_project_scene", sc))
            C.memmove(addr, data.ctypes.data, nbytes)
            ring.publish(k)
        if consumer is not None:
            consumer.join()
            assert order == list(range(nframes))
            np.save(out_path, np.stack(frames))
        dist.barrier()
        ring.close()
    finally:
        dist.destroy_process_group()


def test_ordered_frame_ring_world3_fewer_slots_than_frames(tmp_path):
    out = str(tmp_path / "ring.npy")
    nframes, world = 11, 3
    mp.spawn(_ring_worker, args=(world, _free_port(), nframes, 3, out), nprocs=world, join=True)   # 3 slots: one per rank
    orc = U.load_oracle()
    sc = S.SceneData(16, 9, S.synthetic_cubemap("colors", 16))
    got = np.load(out)
    assert got.shape[0] == nframes
    for k, t in enumerate(sharding.orbit_times(nframes)):
        sc.set_time(t)
        assert np.array_equal(got[k], U.oracle_stream(orc, U.cpu_render(orc, "orc_project_scene", sc))), k


def test_ordered_frame_ring_stop_releases_producers():
    from terminalraytracer_b200 import pipeline
    ring = pipeline.OrderedFrameRing(64, 8, 2)
    assert ring.acquire(0) == ring.slot_address(0) and ring.acquire(1) == ring.slot_address(1)
    ring.publish(0)
    seen = []
    assert ring.consume(lambda k, v: seen.append(k) or True) == 1      # the writer asks to stop after frame 0
    assert seen == [0] and ring.acquire(2) is None                       # producers are released with "stop"
    ring.close()
