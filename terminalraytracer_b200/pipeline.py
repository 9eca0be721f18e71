"""Frame pipelines on top of the C ABI: K1 (render band -> quantised cells) -> K2 (encode band ->
bytes) -> gather on rank 0.  torch owns the device buffers and the process group; the kernels are
libtrt_b200's and run on torch's current stream (Renderer.use_stream), so CUDA events recorded by the
caller bracket them."""
import torch

from . import abi, dist as tdist, sharding


class _DevicePointer:
    """__cuda_array_interface__ view of `nbytes` bytes at a raw device pointer (memory owned by libtrt_b200)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class SharedHostStream:
    """A host buffer for the terminal stream that every rank of the node has mapped and page-locked: POSIX shared memory
    (a file under /dev/shm, unlinked as soon as everybody has it open) + trt_host_register.  FramePipeline(host_stream=
    self.ptr) lets each rank copy its own row bands there; rank 0 then holds the complete stream in `self.array`."""

    @staticmethod
    def available(nbytes, rank=0, world_size=1, group=None):
        """True on every rank when /dev/shm can hold the buffer (rank 0 looks, everybody hears)."""
        import os
        box = [None]
        if rank == 0:
            try:
                st = os.statvfs("/dev/shm")
                box[0] = bool(st.f_bavail * st.f_frsize >= int(nbytes) + (64 << 20))
            except OSError:
                box[0] = False
        if world_size > 1:
            import torch.distributed as dist
            dist.broadcast_object_list(box, src=0, group=group)
        return bool(box[0])

    def __init__(self, renderer, nbytes, rank=0, world_size=1, group=None):
        import ctypes as C
        import mmap
        import os
        import numpy as np
        self.r, self.nbytes = renderer, int(nbytes)
        self.rank, self.world_size = rank, world_size
        self.flags_at = (self.nbytes + 4095) & ~4095          # one int64 per rank behind the stream: the step it has delivered
        self.map_bytes = self.flags_at + 4096
        box = [None]
        if rank == 0:
            box[0] = "/dev/shm/trt_b200_stream_%d_%x" % (os.getpid(), id(self) & 0xffffff)
            fd = os.open(box[0], os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
            os.ftruncate(fd, self.map_bytes)
        if world_size > 1:
            import torch.distributed as dist
            dist.broadcast_object_list(box, src=0, group=group)
            if rank != 0:
                fd = os.open(box[0], os.O_RDWR)
        self.map = mmap.mmap(fd, self.map_bytes)
        os.close(fd)
        if world_size > 1:
            dist.barrier(group=group)
        if rank == 0:
            os.unlink(box[0])                 # the mappings keep the memory alive; nothing is left behind on a crash
        self.array = np.frombuffer(self.map, dtype=np.uint8, count=self.nbytes)
        self.flags = np.frombuffer(self.map, dtype=np.int64, count=max(world_size, 1), offset=self.flags_at)
        self.ptr = C.addressof(C.c_char.from_buffer(self.map))
        self.r.L.trt_host_register(self.ptr, self.nbytes)

    def arrive_and_wait(self, step, timeout_s=60.0):
        """this rank's bytes of frame `step` are in the buffer (its copies have completed); returns when everybody's are —
        a barrier through the shared mapping itself: no collective, no device work"""
        import time
        self.flags[self.rank] = step
        deadline = time.monotonic() + timeout_s
        while int(self.flags.min()) < step:
            if time.monotonic() > deadline:
                raise RuntimeError("SharedHostStream: a rank did not deliver frame %d" % step)
            time.sleep(0)

    def close(self):
        if self.ptr:
            self.r.L.trt_host_unregister(self.ptr)
            self.ptr = None
            self.array = self.flags = None     # the mmap itself is released with the object (ctypes keeps an export on it)


class FramePipeline:
    """One big frame, row-band sharded across `world_size` ranks (BASELINE configs 1-3).

    Exchange (the only one on the path), three forms:
      * `peer=True`: every rank writes its encoded bytes straight into rank 0's DEVICE stream buffer over NVLink peer
        memory (CUDA IPC handle exchanged once through the process group; copy-engine transfers, trt_push_to_peer), piece
        by piece while its next piece renders, and a collective ends the step.
      * `host_stream=<address>`: the terminal stream is wanted in HOST memory (what the caller fwrite()s).  The address is a
        page-locked buffer every rank has mapped (shared memory + trt_host_register): each rank copies its own pieces
        there over its own PCIe link, so the device-to-host transfer is spread over N links instead of funnelled through
        rank 0's, and no device-side gather is needed at all.
      * neither: the bands are gathered with NCCL/gloo send-recv after the render (dist.gather_bands; this is what the CPU
        tests exercise).
    `fused=True` (with peer or host_stream): no separate encode kernel and no copies at all — K1 encodes every finished tile
    and stores its bytes directly at their place in the destination stream (trt_render_rows_ansi_device), rank 0's memory
    over NVLink or the shared host buffer over PCIe: one kernel launch per rank and frame, the transfer rides along tile by
    tile.
    `direct=True` (with peer): K1 writes quantised cells locally and K2 — one launch per piece of the band — stores the encoded
    bytes straight into rank 0's stream through the peer mapping (its bulk copies target the peer address); no copy engine, no K1
    epilogue.
    `adapt=True` (with peer or host_stream): the collective that ends a step carries every rank's measured K1 time, and the
    bands of the next step follow from it (sharding.reweight): the picture changes slowly from frame to frame, so after a
    few frames the ranks finish together.  Bands only decide who renders which rows: the stream is byte-identical for any
    split."""

    def __init__(self, renderer, width, height, rank=0, world_size=1, row_weights=None, group=None, peer=False, pieces=(0.7, 0.3),
                 adapt=False, host_stream=None, fused=False, host_sync=None, direct=False):
        self.r = renderer
        self.direct = bool(direct)           # peer: K2 itself stores into rank 0's stream through the peer mapping (no push)
        self.host_sync = host_sync           # SharedHostStream whose arrival flags end a host_stream step (else: a collective)
        self.step_no = 0
        self.k1_times = None
        self.settled = False                 # adapt: the last feedback step saw the slowest rank within 2 % of the mean
        self.width, self.height = width, height
        self.rank, self.world_size, self.group = rank, world_size, group
        self.device = torch.device("cuda", renderer.device)
        self.host_stream = int(host_stream) if host_stream else None
        self.peer = bool(peer) and world_size > 1 and not self.host_stream
        self.async_pieces = self.peer or bool(self.host_stream)
        self.adapt = bool(adapt) and world_size > 1 and self.async_pieces
        self.fused = bool(fused) and self.async_pieces
        self.direct = self.direct and self.peer and not self.fused
        self.piece_fractions = pieces if (self.async_pieces and not self.fused) else 1
        self.weights = None if row_weights is None else [float(x) for x in row_weights]
        if self.adapt and self.weights is None:
            self.weights = [1.0] * height
        self._set_bands(sharding.row_bands(height, world_size, self.weights))
        # with adaptive bands any rank may come to own any row: local buffers cover the frame and are indexed by row
        self.base_row = 0 if self.adapt else self.row0
        rows = height if self.adapt else self.row1 - self.row0
        if self.fused:
            rows = 0                              # K1 stores the encoded bytes at their destination: no intermediate buffers
        self.quant = torch.empty(max(rows * width, 1) * 4, dtype=torch.uint8, device=self.device)
        self.stream_ptr = self.peer_base = None
        self.k1_span, self.k1_launches = None, 0
        self.flags_offset = (abi.stream_bytes(width, height) + 16 + 255) & ~255
        self.rebalance_above = 1.005          # adapt: move the cuts when the slowest rank is this far above the mean
        total = abi.stream_bytes(width, height)
        if rank == 0 and not self.host_stream:
            if self.peer:
                # cudaMalloc'ed by the library (an IPC handle needs the base of an allocation), viewed as a torch tensor
                # + 128 flag words behind the stream (trt_signal_step / trt_wait_steps): word r = rank r's completed step,
                # word 32 = a wait timed out, word 40 = the last frame rank 0 has consumed
                self.flags_offset = (total + 16 + 255) & ~255
                self.stream_ptr = self.r.L.trt_device_alloc(self.flags_offset + 512)
                self.stream = torch.as_tensor(_DevicePointer(self.stream_ptr, total), device=self.device)
                torch.as_tensor(_DevicePointer(self.stream_ptr + self.flags_offset, 512), device=self.device).zero_()
                torch.cuda.synchronize(self.device)
            else:
                self.stream = torch.empty(total, dtype=torch.uint8, device=self.device)
            self.band_bytes = None
        else:
            self.stream = None
            self.band_bytes = torch.empty(max(rows * abi.row_bytes(width), 1), dtype=torch.uint8, device=self.device)
        if self.host_stream and rank == 0:
            import ctypes as C
            C.memmove(self.host_stream, abi.HOME, abi.HOME_BYTES)               # ESC[0;0H, TRT.c:1144
            C.memset(self.host_stream + total - abi.TAIL_NULS, 0, abi.TAIL_NULS)
        if self.peer:
            import ctypes as C
            import torch.distributed as dist
            box = [None]
            if rank == 0:
                handle = (C.c_ubyte * 64)()
                self.r.L.trt_ipc_export(self.stream_ptr, handle)
                box[0] = bytes(handle)
            dist.broadcast_object_list(box, src=0, group=group)
            if rank != 0:
                handle = (C.c_ubyte * 64).from_buffer_copy(box[0])
                self.peer_base = self.r.L.trt_ipc_import(handle)
        self.r.use_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def _set_bands(self, bands):
        self.bands = bands
        self.row0, self.row1 = bands[self.rank]
        self.pieces = sharding.sub_bands(bands[self.rank], self.piece_fractions, self.weights)

    def close(self):
        if self.peer_base:
            self.r.L.trt_ipc_close(self.peer_base)
            self.peer_base = None
        if self.stream_ptr:
            self.stream = None
            self.r.L.trt_device_free(self.stream_ptr)
            self.stream_ptr = None

    def render_local(self, scene, k1_events=None):
        """K1 + K2 for this rank's band, piece by piece (asynchronous); with peer=True / host_stream every finished piece
        is sent on its way while the next one renders.  k1_events: optional list receiving one (start, end) pair of torch
        events per call: start of the first K1 launch, end of the last."""
        rb = abi.row_bytes(self.width)
        self.r.set_scene_async(scene)        # ordered on the stream like the kernels behind it: no host wait per frame
        if self.stream is not None:
            self.r.stream_frame(self.stream.data_ptr(), self.width, self.height)
        # (Pieces on two alternating streams — the next piece's K1 filling the SMs that the tail of the previous one leaves
        # idle — were measured: the tails are short, ~0.04 ms, and the persistent K1 of the next piece then keeps the
        # previous piece's K2 and with it its transfer off the SMs until it ends.  One stream it is.)
        timed = (self.adapt and self._feedback_due_next()) or k1_events is not None
        first = last = None
        remote = self.peer and self.stream is None           # this rank writes into rank 0's memory
        if remote and self.step_no > 0 and not self.direct:
            # back-pressure: rank 0 must have taken frame step_no before its bytes are overwritten; fused: the kernel itself
            # stores there, so it waits; pieces: only the pushes wait (on the copy stream), K1 and K2 run ahead; direct: K2 waits
            self.r.L.trt_wait_steps(self.peer_base + self.flags_offset + 4 * 40, 1, self.step_no, 0 if self.fused else 1)
        ordered = False
        for i, (r0, r1) in enumerate(self.pieces):
            q = self.quant.data_ptr() + (r0 - self.base_row) * self.width * 4
            if timed and first is None:
                first = torch.cuda.Event(enable_timing=True)
                first.record()
            if self.fused:
                dst = self.host_stream or (self.stream.data_ptr() if self.stream is not None else self.peer_base)
                self.r.render_rows_ansi(self.width, self.height, r0, r1, dst)
            else:
                self.r.render_rows_quant(self.width, self.height, r0, r1, q)
            if timed and i == len(self.pieces) - 1:
                last = torch.cuda.Event(enable_timing=True)
                last.record()
            if self.fused:
                continue
            if self.stream is not None:
                self.r.encode_rows_quant(q, self.width, r1 - r0, self.stream.data_ptr(), abi.HOME_BYTES + r0 * rb)
            elif self.direct:
                # K2's bulk stores (cp.async.bulk shared -> global) go through the peer mapping: the band's bytes cross NVLink
                # as they are encoded, no copy engine, no second pass over them
                # (piece by piece: the first pieces cross while the next one renders; all ranks' last pieces arrive at rank 0
                # together, so the exposed part is the LAST piece's share of the frame over rank 0's NVLink ingest rate)
                if self.step_no > 0 and i == 0:
                    self.r.L.trt_wait_steps(self.peer_base + self.flags_offset + 4 * 40, 1, self.step_no, 0)
                self.r.encode_rows_quant(q, self.width, r1 - r0, self.peer_base, abi.HOME_BYTES + r0 * rb)
            else:
                off = (r0 - self.base_row) * rb
                if self.peer and not ordered:
                    self.r.L.trt_stream_wait_copies()       # the previous step's pushes still read band_bytes
                    ordered = True
                self.r.encode_rows_quant(q, self.width, r1 - r0, self.band_bytes.data_ptr(), off)
                dst = self.host_stream if self.host_stream else self.peer_base
                if dst:
                    self.r.L.trt_push_to_peer(dst + abi.HOME_BYTES + r0 * rb, self.band_bytes.data_ptr() + off, (r1 - r0) * rb)
        # K1 time of this rank and frame: from the start of the first piece to the end of the last one (includes the small K2s between)
        self.k1_span = (first, last) if (first is not None and last is not None) else None
        if k1_events is not None and self.k1_span:
            k1_events.append(self.k1_span)
        self.k1_launches += len(self.pieces)

    FEEDBACK_STEPS = 5        # adapt: the first steps of a pipeline exchange their K1 times every step ...
    FEEDBACK_UNTIL = 16       # ... and go on doing so while the slowest rank is more than 2 % above the mean, at most this long;
    FEEDBACK_EVERY = 32       # later only every so many steps

    def _feedback_at(self, step):
        return step <= self.FEEDBACK_STEPS or (not self.settled and step <= self.FEEDBACK_UNTIL) or step % self.FEEDBACK_EVERY == 0

    def _feedback_due_next(self):
        return self._feedback_at(self.step_no + 1)

    def _feedback_due(self):
        """adapt: measure and exchange the K1 times on this step?  Every one of the first five steps, on while the slowest rank is
        more than 2 % above the mean (at most 16 steps), then every 32nd — a slowly changing picture keeps its balance, and a step
        without feedback needs no host synchronisation at all.  (Round 2's first policy, "until the ranks are level within 0.5 % three steps running", never
        stopped at 8 GPUs — row granularity and timing noise leave 1-1.5 % — and its per-step synchronisation cost 0.25 ms of a
        3.4 ms step.)"""
        return self._feedback_at(self.step_no)

    def gather(self):
        self.step_no += 1
        if self.async_pieces:
            import torch.distributed as dist
            feedback = self.adapt and self._feedback_due()
            if self.peer:
                # completion on the device: this rank's step number lands in rank 0's flag array behind its bytes, rank 0's
                # stream waits for all of them; no host takes part
                base = self.stream_ptr if self.stream_ptr else self.peer_base
                self.r.L.trt_signal_step(base + self.flags_offset + 4 * self.rank, self.step_no,
                                         0 if (self.fused or self.direct or self.stream is not None) else 1)
                if self.rank == 0:
                    self.r.L.trt_wait_steps(self.stream_ptr + self.flags_offset, self.world_size, self.step_no, 0)
                    # (a consumer of the finished frame — a device-to-host copy, a display — is enqueued here, on this stream)
                    self.r.L.trt_signal_step(self.stream_ptr + self.flags_offset + 4 * 40, self.step_no, 0)
            else:
                # host stream: the bytes must be in host memory when the step ends — wait for this rank's copies (or, fused, for
                # the kernel that stored them), then meet the others through the flags of the shared mapping
                if self.fused:
                    torch.cuda.current_stream(self.device).synchronize()
                else:
                    self.r.L.trt_peer_copies_wait()
                if self.host_sync is not None:
                    self.host_sync.arrive_and_wait(self.step_no)
                elif self.world_size > 1 and not feedback:
                    dist.barrier(group=self.group)
            if not feedback:
                return self.stream
            # feedback step: every rank's K1 time of this frame through one small collective
            if self.k1_span:
                self.k1_span[1].synchronize()
            where = self.device if dist.get_backend(self.group) == "nccl" else "cpu"
            mine = torch.tensor([self.k1_span[0].elapsed_time(self.k1_span[1]) if self.k1_span else 0.0], dtype=torch.float32, device=where)
            times = torch.empty(self.world_size, dtype=torch.float32, device=where)
            dist.all_gather_into_tensor(times, mine, group=self.group)
            self.k1_times = times.tolist()
            # every rank sees the same numbers and takes the same decision; bands that are already level are left alone
            busy = [t for t in self.k1_times if t > 0]
            self.settled = bool(busy) and max(busy) * len(busy) <= 1.02 * sum(busy)
            if busy and max(busy) * len(busy) > self.rebalance_above * sum(busy):
                self.weights = sharding.reweight(self.weights, self.bands, self.k1_times)
                self._set_bands(sharding.row_bands(self.height, self.world_size, self.weights))
            return self.stream
        band = None
        if self.rank != 0:
            n = (self.row1 - self.row0) * abi.row_bytes(self.width)
            band = self.band_bytes[:n]
        return tdist.gather_bands(self.stream, band, self.width, self.bands, self.rank, self.world_size, self.group)

    def finish(self):
        """after the last step: the stream is complete on rank 0 (device-side completion leaves the hosts un-synchronised)"""
        torch.cuda.synchronize(self.device)
        if self.peer:
            self.r.L.trt_peer_copies_wait()
            if self.world_size > 1:
                import torch.distributed as dist
                dist.barrier(group=self.group)
            torch.cuda.synchronize(self.device)
            if self.rank == 0:
                flags = torch.as_tensor(_DevicePointer(self.stream_ptr + self.flags_offset, 512), device=self.device).view(torch.int32).cpu()
                if int(flags[32]) != 0 or int(flags[72]) != 0:
                    raise RuntimeError("FramePipeline: a rank never signalled its step (trt_wait_steps timed out)")
        return self.stream

    def render(self, scene, k1_events=None):
        """Full step: returns the complete byte stream (device tensor) on rank 0, None elsewhere (and None everywhere with
        host_stream: the bytes are in the host buffer)."""
        self.render_local(scene, k1_events)
        out = self.gather()
        if self.peer:
            self.finish()        # callers of render() read the stream right away
        return out


class OrderedFrameRing:
    """The ordered sink of a frame-sharded animation: a ring of `slots` frame buffers in POSIX shared memory that every rank
    of the node maps (and page-locks, when a renderer is given, so that the device-to-host copies land in it directly), a
    ready flag per frame and the count of frames consumed.  Rank r produces frames r, r+N, ... (acquire -> the frame's bytes
    arrive -> publish); ONE consumer takes the frames strictly in order while later ones are still rendering.  Frame k lives
    in slot k mod slots; a producer may run at most `slots` frames ahead of the consumer.  Host logic only (mmap + numpy):
    the CPU tests drive it with plain processes."""

    HEADER = 4096

    def __init__(self, frame_bytes, n_frames, slots, rank=0, world_size=1, group=None, renderer=None):
        import ctypes as C
        import mmap
        import os
        import numpy as np
        self.frame_bytes, self.n_frames, self.slots = int(frame_bytes), int(n_frames), int(slots)
        self.rank, self.world_size, self.r = rank, world_size, renderer
        self.slot_stride = (self.frame_bytes + 4095) & ~4095
        self.flags_bytes = (4 * self.n_frames + 4095) & ~4095
        self.nbytes = self.HEADER + self.flags_bytes + self.slots * self.slot_stride
        box = [None]
        if rank == 0:
            box[0] = "/dev/shm/trt_b200_ring_%d_%x" % (os.getpid(), id(self) & 0xffffff)
            fd = os.open(box[0], os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
            os.ftruncate(fd, self.nbytes)          # zero-filled: nothing ready, nothing consumed
        if world_size > 1:
            import torch.distributed as dist
            dist.broadcast_object_list(box, src=0, group=group)
            if rank != 0:
                fd = os.open(box[0], os.O_RDWR)
        self.map = mmap.mmap(fd, self.nbytes)
        os.close(fd)
        if world_size > 1:
            dist.barrier(group=group)
        if rank == 0:
            os.unlink(box[0])
        self.base = C.addressof(C.c_char.from_buffer(self.map))
        self.header = np.frombuffer(self.map, dtype=np.int64, count=2, offset=0)          # [consumed, stop]
        self.ready = np.frombuffer(self.map, dtype=np.int32, count=self.n_frames, offset=self.HEADER)
        self.slots_base = self.base + self.HEADER + self.flags_bytes
        self.registered = False
        if renderer is not None:
            renderer.L.trt_host_register(self.slots_base, self.slots * self.slot_stride)
            self.registered = True

    # ---- producers -------------------------------------------------------------------------------------------------
    def slot_address(self, frame):
        return self.slots_base + (frame % self.slots) * self.slot_stride

    def acquire(self, frame, poll_s=20e-6):
        """address of frame's slot once the consumer has freed it (frame - slots consumed); None after stop()"""
        import time
        while frame >= int(self.header[0]) + self.slots:
            if self.header[1]:
                return None
            time.sleep(poll_s)
        return None if self.header[1] else self.slot_address(frame)

    def publish(self, frame):
        self.ready[frame] = 1

    def stop(self):
        self.header[1] = 1

    # ---- the one consumer ------------------------------------------------------------------------------------------
    def reset(self, n_frames=None):
        """before a new animation (one process, between barriers): nothing ready, nothing consumed"""
        self.ready[:self.n_frames if n_frames is None else n_frames] = 0
        self.header[0] = 0
        self.header[1] = 0

    def consume(self, write, n_frames=None, poll_s=20e-6, timeout_s=600.0):
        """frames 0..n_frames-1 in order: write(frame_index, memoryview of its bytes) as soon as frame k is ready; returns the
        number of frames written (fewer than n_frames after stop() or when write returns a true value)"""
        import time
        view = memoryview(self.map)
        off0 = self.HEADER + self.flags_bytes
        deadline = time.monotonic() + timeout_s
        k = 0
        total = self.n_frames if n_frames is None else n_frames
        try:
            while k < total:
                while not self.ready[k]:
                    if self.header[1] or time.monotonic() > deadline:
                        return k
                    time.sleep(poll_s)
                off = off0 + (k % self.slots) * self.slot_stride
                halt = write(k, view[off:off + self.frame_bytes])
                k += 1
                self.header[0] = k
                if halt:
                    self.stop()
                    break
        finally:
            view.release()
        return k

    def close(self):
        if self.registered:
            self.r.L.trt_host_unregister(self.slots_base)
            self.registered = False
        self.header = self.ready = None


class OrbitPipeline:
    """Animation, frame-sharded with streamed, ordered output (BASELINE config 4; the reference's frame loop
    TRT.c:1317-1367 with one fwrite per frame, TRT.c:1171): frame k is rendered by rank k mod N through
    trt_render_orbit_to, its bytes go from the device straight into a shared page-locked ring (OrderedFrameRing) over the
    rank's own PCIe link, and rank 0 writes the frames out strictly in order while later ones render.  Nothing is gathered
    at the end and no rank ever holds more than `slots_per_rank` finished frames."""

    def __init__(self, renderer, width, height, rank=0, world_size=1, group=None, slots_per_rank=3, max_frames=4096):
        self.r = renderer
        self.width, self.height = width, height
        self.rank, self.world_size, self.group = rank, world_size, group
        self.slots_per_rank = slots_per_rank
        self.frame_bytes = abi.stream_bytes(width, height)
        # mapped and page-locked once (registering a gigabyte takes longer than rendering it); every stream() reuses it
        self.ring = OrderedFrameRing(self.frame_bytes, max_frames, slots_per_rank * world_size, rank, world_size, group, renderer)

    def close(self):
        if self.ring is not None:
            self.ring.close()
            self.ring = None

    def _barrier(self):
        if self.world_size > 1:
            import torch.distributed as dist
            dist.barrier(group=self.group)

    def stream(self, scene, times, write=None):
        """Render this rank's frames of the camera path `times` (scene: un-posed camera, as SceneData builds it); rank 0 also
        consumes: write(frame_index, memoryview) for k = 0, 1, 2, ... in order (default: discard).  Returns the number of
        frames this rank rendered and — on rank 0 — the number written out."""
        import ctypes as C
        import threading
        from . import lib as _lib
        n = len(times)
        ring = self.ring
        if n > ring.n_frames:
            raise ValueError("OrbitPipeline: %d frames, ring sized for %d (max_frames)" % (n, ring.n_frames))
        self._barrier()                          # nobody is still inside the previous stream()
        if self.rank == 0:
            ring.reset(n)
        self._barrier()
        written = [0]
        consumer = None
        if self.rank == 0:
            sink = write if write is not None else (lambda k, view: False)
            consumer = threading.Thread(target=lambda: written.__setitem__(0, ring.consume(sink, n)), daemon=True)
            consumer.start()
        arr = (C.c_double * n)(*[float(t) for t in times])

        def _acquire(frame, nbytes, _user):
            return ring.acquire(frame)          # None -> NULL: the loop stops

        def _landed(ptr, nbytes, frame, _user):
            ring.publish(frame)
            return 0

        acq, snk = _lib.FRAME_ACQUIRE(_acquire), _lib.FRAME_SINK(_landed)
        self.r.use_stream(None)                 # the orbit loop runs on the library's own streams
        done = self.r.L.trt_render_orbit_to(C.byref(scene.c), self.width, self.height, arr, n, self.rank, self.world_size,
                                            C.cast(acq, C.c_void_p), C.cast(snk, C.c_void_p), None)
        if consumer is not None:
            consumer.join()
        return done, written[0]

    def collect(self, scene, times):
        """small animations (tests): the ordered frames as a list of uint8 arrays on rank 0, None elsewhere"""
        import numpy as np
        frames = []
        self.stream(scene, times, (lambda k, view: frames.append(np.frombuffer(view, dtype=np.uint8).copy()) and False))
        return frames if self.rank == 0 else None
