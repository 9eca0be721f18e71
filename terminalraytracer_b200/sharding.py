"""How the render path is split across GPUs: contiguous row bands of one frame, or
whole frames of an animation by frame index.  Pure host logic, no CUDA, no torch."""
from . import abi


def row_bands(height, world_size, weights=None):
    """Split rows [0,height) into world_size contiguous bands [(r0,r1), ...].

    weights: optional per-row cost estimates (len == height); bands then carry roughly equal cost
    (sky rows are ~5x cheaper than sphere/ground rows).  Bands stay contiguous, so
    the gathered byte stream is a plain concatenation.  Every row belongs to exactly one band; bands
    may be empty when world_size > height."""
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    if weights is None:
        base, extra = divmod(height, world_size)
        bands, r = [], 0
        for i in range(world_size):
            n = base + (1 if i < extra else 0)
            bands.append((r, r + n))
            r += n
        return bands
    if len(weights) != height:
        raise ValueError("need one weight per row")
    total = float(sum(weights))
    if total <= 0:
        return row_bands(height, world_size)
    bands, r, acc = [], 0, 0.0
    for i in range(world_size):
        target = total * (i + 1) / world_size
        r1 = r
        while r1 < height and (acc + weights[r1] <= target or i == world_size - 1):
            acc += weights[r1]
            r1 += 1
        # always leave enough rows for nobody to be forced negative; allow empty bands
        bands.append((r, r1))
        r = r1
    bands[-1] = (bands[-1][0], height)
    return bands


def sub_bands(band, pieces, weights=None):
    """Split one rank's band (r0,r1) into contiguous, non-empty pieces: each piece is pushed to rank 0 while the next one
    renders (pipeline.FramePipeline), so only the last piece's transfer is exposed.  pieces: an int (equal cost) or a
    sequence of cost fractions, e.g. (0.7, 0.3) — a big first piece and a small last one keep both the number of
    kernel launches and the exposed transfer small."""
    r0, r1 = band
    if r1 <= r0:
        return []
    n = r1 - r0
    w = [1.0] * n if weights is None else [float(x) for x in weights[r0:r1]]
    fractions = [1.0 / int(pieces)] * int(pieces) if isinstance(pieces, int) else [float(f) for f in pieces]
    total, norm = sum(w), sum(fractions)
    if total <= 0 or norm <= 0:
        return [(r0, r1)]
    out, start, acc, cut = [], 0, 0.0, 0.0
    for k, f in enumerate(fractions):
        cut += f / norm
        end = start
        if k == len(fractions) - 1:
            end = n
        else:
            while end < n and acc + w[end] <= total * cut:
                acc += w[end]
                end += 1
        if end > start:
            out.append((r0 + start, r0 + end))
            start = end
    if start < n:
        out.append((r0 + start, r1))
    return out


def band_byte_range(width, band):
    """Byte range of a row band inside the terminal stream (home sequence included in the offset)."""
    r0, r1 = band
    rb = abi.row_bytes(width)
    return abi.HOME_BYTES + r0 * rb, abi.HOME_BYTES + r1 * rb


def frames_for_rank(num_frames, rank, world_size):
    """Animation sharding: frame k is rendered by rank k mod world_size (BASELINE config 4)."""
    return list(range(rank, num_frames, world_size))


def orbit_times(num_frames, period_s=20.0):
    """t_k for a full yaw turn at the reference's 0.05 rev/s (TRT.c:1332): k * period / n."""
    return [k * (period_s / num_frames) for k in range(num_frames)]
