"""How the render path is split across GPUs: contiguous row bands of one frame, or
whole frames of an animation by frame index.  Pure host logic, no CUDA, no torch."""
import numpy as np

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
    # running sums in row order (np.cumsum adds sequentially, so every rank gets the same doubles and the same cuts)
    cs = np.cumsum(np.asarray(weights, dtype=np.float64))
    total = float(cs[-1]) if height else 0.0
    if total <= 0:
        return row_bands(height, world_size)
    bands, r = [], 0
    for i in range(world_size):
        if i == world_size - 1:
            r1 = height
        else:
            # the rows whose running sum stays within this band's share of the total cost
            r1 = max(r, int(np.searchsorted(cs, total * (i + 1) / world_size, side="right")))
        bands.append((r, r1))   # empty bands are allowed
        r = r1
    return bands


def reweight(weights, bands, times):
    """Feedback for the next frame: scale the per-row cost estimates inside every band so that the band's sum equals the
    time its rank measured for it (K1 milliseconds).  Bands derived from the result (row_bands) move rows from the slow
    ranks to the fast ones; with a slowly changing picture — an orbit, or the same frame again — a few frames are enough
    for the ranks to finish together.  Bands without a usable measurement keep their estimates (scaled like the rest)."""
    w = np.array(weights, dtype=np.float64)
    measured = [(r0, r1, float(t)) for (r0, r1), t in zip(bands, times) if r1 > r0 and t > 0 and float(w[r0:r1].sum()) > 0]
    if not measured:
        return w
    # overall scale for the rows nobody measured: time per unit of estimated cost over the measured bands
    scale = sum(t for _, _, t in measured) / sum(float(w[r0:r1].sum()) for r0, r1, _ in measured)
    out = w * scale
    for r0, r1, t in measured:
        out[r0:r1] = w[r0:r1] * (t / float(w[r0:r1].sum()))
    return out


def sub_bands(band, pieces, weights=None):
    """Split one rank's band (r0,r1) into contiguous, non-empty pieces: each piece is pushed to rank 0 while the next one
    renders (pipeline.FramePipeline), so only the last piece's transfer is exposed.  pieces: an int (equal cost) or a
    sequence of cost fractions, e.g. (0.7, 0.3) — a big first piece and a small last one keep both the number of
    kernel launches and the exposed transfer small."""
    r0, r1 = band
    if r1 <= r0:
        return []
    n = r1 - r0
    w = np.ones(n) if weights is None else np.asarray(weights[r0:r1], dtype=np.float64)
    fractions = [1.0 / int(pieces)] * int(pieces) if isinstance(pieces, int) else [float(f) for f in pieces]
    cs = np.cumsum(w)
    total, norm = float(cs[-1]), sum(fractions)
    if total <= 0 or norm <= 0:
        return [(r0, r1)]
    out, start, cut = [], 0, 0.0
    for k, f in enumerate(fractions):
        cut += f / norm
        end = n if k == len(fractions) - 1 else max(start, int(np.searchsorted(cs, total * cut, side="right")))
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
