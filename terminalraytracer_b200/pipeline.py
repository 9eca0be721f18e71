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


class FramePipeline:
    """One big frame, row-band sharded across `world_size` ranks (BASELINE configs 1-3).

    Exchange (the only one on the path): with `peer=True` and world_size > 1 every rank writes its encoded bytes
    straight into rank 0's stream buffer over NVLink peer memory (CUDA IPC handle exchanged once through the process
    group; copy-engine transfers, trt_push_to_peer), piece by piece while its next piece renders, and a barrier ends
    the step.  With `peer=False` the bands are gathered with NCCL/gloo send-recv after the render (dist.gather_bands;
    this is what the CPU tests exercise)."""

    def __init__(self, renderer, width, height, rank=0, world_size=1, row_weights=None, group=None, peer=False, pieces=(0.7, 0.3)):
        self.r = renderer
        self.width, self.height = width, height
        self.rank, self.world_size, self.group = rank, world_size, group
        self.device = torch.device("cuda", renderer.device)
        self.bands = sharding.row_bands(height, world_size, row_weights)
        self.row0, self.row1 = self.bands[rank]
        rows = self.row1 - self.row0
        self.peer = bool(peer) and world_size > 1
        self.pieces = sharding.sub_bands(self.bands[rank], pieces if self.peer else 1, row_weights)
        self.quant = torch.empty(max(rows * width, 1) * 4, dtype=torch.uint8, device=self.device)
        self.stream_ptr = self.peer_base = None
        total = abi.stream_bytes(width, height)
        if rank == 0:
            if self.peer:
                # cudaMalloc'ed by the library (an IPC handle needs the base of an allocation), viewed as a torch tensor
                self.stream_ptr = self.r.L.trt_device_alloc(total + 16)
                self.stream = torch.as_tensor(_DevicePointer(self.stream_ptr, total), device=self.device)
            else:
                self.stream = torch.empty(total, dtype=torch.uint8, device=self.device)
            self.band_bytes = None
        else:
            self.stream = None
            self.band_bytes = torch.empty(max(rows * abi.row_bytes(width), 1), dtype=torch.uint8, device=self.device)
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

    def close(self):
        if self.peer_base:
            self.r.L.trt_ipc_close(self.peer_base)
            self.peer_base = None
        if self.stream_ptr:
            self.stream = None
            self.r.L.trt_device_free(self.stream_ptr)
            self.stream_ptr = None

    def render_local(self, scene, k1_events=None):
        """K1 + K2 for this rank's band, piece by piece (asynchronous); with peer=True every finished piece is pushed
        into rank 0's stream.  k1_events: optional list receiving (start, end) torch events around every K1 launch."""
        rb = abi.row_bytes(self.width)
        self.r.set_scene(scene)
        if self.rank == 0:
            self.r.stream_frame(self.stream.data_ptr(), self.width, self.height)
        for (r0, r1) in self.pieces:
            q = self.quant.data_ptr() + (r0 - self.row0) * self.width * 4
            if k1_events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            self.r.render_rows_quant(self.width, self.height, r0, r1, q)
            if k1_events is not None:
                e1.record()
                k1_events.append((e0, e1))
            if self.rank == 0:
                self.r.encode_rows_quant(q, self.width, r1 - r0, self.stream.data_ptr(), abi.HOME_BYTES + r0 * rb)
            else:
                off = (r0 - self.row0) * rb
                self.r.encode_rows_quant(q, self.width, r1 - r0, self.band_bytes.data_ptr(), off)
                if self.peer:
                    self.r.L.trt_push_to_peer(self.peer_base + abi.HOME_BYTES + r0 * rb, self.band_bytes.data_ptr() + off, (r1 - r0) * rb)

    def gather(self):
        if self.peer:
            import torch.distributed as dist
            if self.rank != 0:
                self.r.L.trt_peer_copies_wait()      # this rank's bytes have landed in rank 0's memory
            dist.barrier(group=self.group)           # ... and so have everybody else's
            return self.stream if self.rank == 0 else None
        band = None
        if self.rank != 0:
            n = (self.row1 - self.row0) * abi.row_bytes(self.width)
            band = self.band_bytes[:n]
        return tdist.gather_bands(self.stream, band, self.width, self.bands, self.rank, self.world_size, self.group)

    def render(self, scene, k1_events=None):
        """Full step: returns the complete byte stream (device tensor) on rank 0, None elsewhere."""
        self.render_local(scene, k1_events)
        return self.gather()


class OrbitPipeline:
    """Animation, frame-sharded: frame k is rendered by rank k mod N (BASELINE config 4)."""

    def __init__(self, renderer, width, height, rank=0, world_size=1, group=None):
        self.frame = FramePipeline(renderer, width, height, 0, 1)  # every rank renders whole frames
        self.r = renderer
        self.width, self.height = width, height
        self.rank, self.world_size, self.group = rank, world_size, group
        self.device = self.frame.device

    def render(self, scene, times):
        ids = sharding.frames_for_rank(len(times), self.rank, self.world_size)
        mine = []
        for k in ids:
            scene.set_time(times[k])
            self.frame.render_local(scene)
            mine.append(self.frame.stream.clone())
        return tdist.gather_frames(mine, ids, len(times), abi.stream_bytes(self.width, self.height), self.rank,
                                   self.world_size, self.device, self.group)
