"""The one exchange step of the path: gathering the per-rank byte bands (or whole
frames) of the terminal stream on rank 0.  One process per GPU, torch.distributed for the plumbing
(NCCL over NVLink on the B200 box, gloo on CPU in the tests).  No other collective exists on this
path: pixels never leave the GPU that rendered them."""
import torch
import torch.distributed as dist

from . import abi, sharding


def gather_bands(stream, band_bytes, width, bands, rank, world_size, group=None):
    """Concatenate the ranks' row-band bytes into `stream` on rank 0.

    stream     : rank 0 only — uint8 tensor of abi.stream_bytes(width, height) bytes whose own band has
                 already been encoded in place; other ranks pass None.
    band_bytes : ranks > 0 — uint8 tensor holding exactly this rank's band bytes (may be empty).
    bands      : [(row0,row1)] * world_size from sharding.row_bands (identical on every rank).
    The stream is row-major with a fixed row size, so rank i's band is the byte range
    sharding.band_byte_range(width, bands[i]) and the gather is a plain concatenation."""
    if world_size == 1:
        return stream
    ops = []
    if rank == 0:
        for src in range(1, world_size):
            b0, b1 = sharding.band_byte_range(width, bands[src])
            if b1 > b0:
                ops.append(dist.P2POp(dist.irecv, stream[b0:b1], src, group))
    else:
        b0, b1 = sharding.band_byte_range(width, bands[rank])
        if b1 > b0:
            assert band_bytes.numel() == b1 - b0, (band_bytes.numel(), b1 - b0)
            ops.append(dist.P2POp(dist.isend, band_bytes, 0, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return stream if rank == 0 else None


def gather_frames(frame_streams, frame_ids, num_frames, stream_bytes, rank, world_size, device, group=None):
    """Animation sharding (frame k on rank k mod N): collect every frame's byte stream on rank 0, in
    frame order.  frame_streams: this rank's list of uint8 tensors, one per entry of frame_ids.
    Returns the list of num_frames tensors on rank 0, None elsewhere."""
    if world_size == 1:
        return list(frame_streams)
    mine = dict(zip(frame_ids, frame_streams))
    out = [None] * num_frames if rank == 0 else None
    ops = []
    for k in range(num_frames):
        owner = k % world_size
        if rank == 0:
            if owner == 0:
                out[k] = mine[k]
            else:
                out[k] = torch.empty(stream_bytes, dtype=torch.uint8, device=device)
                ops.append(dist.P2POp(dist.irecv, out[k], owner, group))
        elif owner == rank:
            ops.append(dist.P2POp(dist.isend, mine[k], 0, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out
