"""Frame pipelines on top of the C ABI: K1 (render band -> quantised cells) -> K2 (encode band ->
bytes) -> gather on rank 0.  torch owns the device buffers and the process group; the kernels are
libtrt_b200's and run on torch's current stream (Renderer.use_stream), so CUDA events recorded by the
caller bracket them."""
import torch

from . import abi, dist as tdist, sharding


class FramePipeline:
    """One big frame, row-band sharded across `world_size` ranks (BASELINE configs 1-3)."""

    def __init__(self, renderer, width, height, rank=0, world_size=1, row_weights=None, group=None):
        self.r = renderer
        self.width, self.height = width, height
        self.rank, self.world_size, self.group = rank, world_size, group
        self.device = torch.device("cuda", renderer.device)
        self.bands = sharding.row_bands(height, world_size, row_weights)
        self.row0, self.row1 = self.bands[rank]
        rows = self.row1 - self.row0
        self.quant = torch.empty(max(rows * width, 1) * 4, dtype=torch.uint8, device=self.device)
        if rank == 0:
            self.stream = torch.empty(abi.stream_bytes(width, height), dtype=torch.uint8, device=self.device)
            self.band_bytes = None
        else:
            self.stream = None
            self.band_bytes = torch.empty(max(rows * abi.row_bytes(width), 1), dtype=torch.uint8, device=self.device)
        self.r.use_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def render_local(self, scene):
        """K1 + K2 for this rank's band (asynchronous)."""
        rows = self.row1 - self.row0
        self.r.set_scene(scene)
        if rows > 0:
            self.r.render_rows_quant(self.width, self.height, self.row0, self.row1, self.quant.data_ptr())
        if self.rank == 0:
            self.r.stream_frame(self.stream.data_ptr(), self.width, self.height)
            if rows > 0:
                b0, _ = sharding.band_byte_range(self.width, self.bands[0])
                self.r.encode_rows_quant(self.quant.data_ptr(), self.width, rows, self.stream.data_ptr(), b0)
        elif rows > 0:
            self.r.encode_rows_quant(self.quant.data_ptr(), self.width, rows, self.band_bytes.data_ptr(), 0)

    def gather(self):
        band = None
        if self.rank != 0:
            n = (self.row1 - self.row0) * abi.row_bytes(self.width)
            band = self.band_bytes[:n]
        return tdist.gather_bands(self.stream, band, self.width, self.bands, self.rank, self.world_size, self.group)

    def render(self, scene):
        """Full step: returns the complete byte stream (device tensor) on rank 0, None elsewhere."""
        self.render_local(scene)
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
