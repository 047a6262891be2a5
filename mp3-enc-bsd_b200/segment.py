"""Single long stream over several GPUs (BASELINE configs[4]): cut at frame boundaries, encode the
segments as independent streams, gather the byte streams on rank 0.

Why this is legal: every stage of the hot path has a bounded memory of the signal — 480 samples (filterbank),
one granule (MDCT overlap), two granules (psy prediction, pre-echo control, block-type state) — so a segment that
is fed `preroll_frames` >= 2 frames of the preceding audio first (outputs discarded) reproduces the spectrum,
thresholds and block types of the whole-stream encode exactly.  One more state has data-dependent memory:
calc_scfsi's statics en[]/xm[] are refreshed only by long-block granules (loop.c:649-667) but read for every
granule 1, so a value can be as old as the current run of short blocks; the default pre-roll of 8 frames covers
runs of up to 16 granules (measured: with 2 frames, 2 of 84 frames of the 320 kbps test stream differ, with 8
none).  The ONE truly unbounded state is the bit reservoir (reservoir.c:101-145).  It is cut at the seam: `Encoder.begin_segment()` empties it, so a segment's first frame has
main_data_begin = 0 and nothing of a segment lies in the previous one; the bytes the previous segment had saved are
left as zero padding in its last frames (valid Layer III: decoders skip them through main_data_begin).  Frames
after a seam therefore differ from the whole-stream encode until the reservoir recurrence re-converges; the
identical-frame fraction is what `frame_identity` reports.  No collective runs on the data path; the only exchange
is one gather of the finished byte streams (NCCL on GPUs, gloo in the CPU tests).

Host-side logic only (the reference has no counterpart: it encodes one stream per process, musicin.c:585).
"""
from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Segment:
    index: int        # position in the stream
    first_frame: int  # first frame whose bytes this segment delivers
    n_frames: int     # frames delivered
    preroll: int      # frames of preceding audio encoded first and discarded (0 for the first segment)


def plan_segments(n_frames, n_segments, preroll_frames=8):
    """balanced contiguous cut of [0, n_frames) into at most n_segments non-empty segments"""
    if n_frames < 1 or n_segments < 1 or preroll_frames < 0:
        raise ValueError("n_frames, n_segments must be >= 1 and preroll_frames >= 0")
    n_segments = min(n_segments, n_frames)
    base, rem = divmod(n_frames, n_segments)
    segs, f = [], 0
    for i in range(n_segments):
        n = base + (1 if i < rem else 0)
        segs.append(Segment(i, f, n, min(preroll_frames, f)))
        f += n
    return segs


def shard_segments(segments, rank, world):
    """contiguous, balanced block of the segment list for one rank (may be empty when world > len(segments))"""
    base, rem = divmod(len(segments), world)
    lo = rank * base + min(rank, rem)
    return segments[lo:lo + base + (1 if rank < rem else 0)]


def segment_pcm(pcm, seg):
    """PCM of one segment incl. its pre-roll, zero-padded to whole frames: int16 [n_ch][(preroll+n_frames)*1152]"""
    n_ch, n = pcm.shape
    a, b = (seg.first_frame - seg.preroll) * 1152, (seg.first_frame + seg.n_frames) * 1152
    out = np.zeros((n_ch, b - a), np.int16)
    hi = min(b, n)
    if hi > a:
        out[:, :hi - a] = pcm[:, a:hi]
    return out


def encode_segments(pcm, segments, encode_batch):
    """Encode `segments` of the stream `pcm` ([n_ch][n] int16).  Segments of equal (n_frames, preroll) are batched:
    `encode_batch(pcm_batch [S][n_ch][(preroll+n_frames)*1152], preroll) -> (uint8 [S][n_frames*FB], lengths [S])`.
    Returns {segment index: (bytes of exactly n_frames*FB, stream length reported by the encoder)}."""
    groups = {}
    for seg in segments:
        groups.setdefault((seg.n_frames, seg.preroll), []).append(seg)
    out = {}
    for (n_frames, preroll), segs in sorted(groups.items()):
        batch = np.stack([segment_pcm(pcm, s) for s in segs])
        data, lengths = encode_batch(batch, preroll)
        for i, s in enumerate(segs):
            out[s.index] = (np.ascontiguousarray(data[i]).tobytes(), int(lengths[i]))
    return out


def gpu_batch_encoder(pkg, sfreq, n_ch, bitrate, device=0, chunk_frames=32):
    """`encode_batch` backed by libmp3gpu.so on one GPU (raises if the library or the GPU is missing: no CPU fallback)"""
    def encode_batch(batch, preroll):
        S, _, n = batch.shape
        F = n // 1152 - preroll
        enc = pkg.Encoder(sfreq, n_ch, bitrate, max_streams=S, max_frames=max(1, min(chunk_frames, max(F, preroll))), device=device)
        step = enc.cfg.max_frames
        for f0 in range(0, preroll, step):                      # pre-roll: run the hot path, discard the outputs
            f1 = min(preroll, f0 + step)
            enc.encode_frames(np.ascontiguousarray(batch[:, :, f0 * 1152:f1 * 1152]), out={})
        enc.begin_segment()
        mp3 = np.zeros((S, F * enc.frame_bytes), np.uint8)
        for f0 in range(0, F, step):
            f1 = min(F, f0 + step)
            enc.encode_frames_mp3(np.ascontiguousarray(batch[:, :, (preroll + f0) * 1152:(preroll + f1) * 1152]), mp3)
        lengths = enc.flush_mp3(mp3, S)
        enc.close()
        return mp3, lengths
    return encode_batch


def gather_payloads(local, n_total, frame_bytes, max_frames, device=None):
    """one gather of the finished segments to rank 0.  `local` = {segment index: (bytes, length)}.  Uses
    torch.distributed when initialised (all_gather of fixed-size rows: index, length, payload), otherwise returns
    `local`.  Returns the full dict on rank 0 and None on the other ranks."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(local)
    world, rank = dist.get_world_size(), dist.get_rank()
    per_rank = (n_total + world - 1) // world
    row = max_frames * frame_bytes
    dev = device if device is not None else "cpu"
    meta = torch.full((per_rank, 3), -1, dtype=torch.int64)
    data = torch.zeros((per_rank, row), dtype=torch.uint8)
    for i, (idx, (payload, length)) in enumerate(sorted(local.items())):
        meta[i, 0], meta[i, 1], meta[i, 2] = idx, length, len(payload)
        data[i, :len(payload)] = torch.frombuffer(bytearray(payload), dtype=torch.uint8)
    meta, data = meta.to(dev), data.to(dev)
    metas = [torch.empty_like(meta) for _ in range(world)]
    datas = [torch.empty_like(data) for _ in range(world)]
    dist.all_gather(metas, meta)
    dist.all_gather(datas, data)
    if rank != 0:
        return None
    out = {}
    for m, d in zip(metas, datas):
        m, d = m.cpu(), d.cpu().numpy()
        for i in range(per_rank):
            idx, length, nbytes = (int(x) for x in m[i])
            if idx >= 0:
                out[idx] = (d[i, :nbytes].tobytes(), length)
    return out


def stitch(segments, payloads):
    """concatenate the segments' byte streams; the last one is cut where the encoder said the stream ends
    (BF_FlushBitstream leaves the final frame short by the reservoir, formatBitstream.c:87-125)"""
    parts = []
    for seg in segments:
        payload, length = payloads[seg.index]
        parts.append(payload[:length] if seg.index == segments[-1].index else payload)
    return b"".join(parts)


def encode_long_stream(pcm, n_segments, encode_batch, frame_bytes, preroll_frames=8, rank=0, world=1, device=None):
    """The whole configs[4] path: plan, shard over ranks, encode the local segments, gather, stitch.
    Returns the byte stream on rank 0, None elsewhere."""
    n_frames = (pcm.shape[1] + 1151) // 1152
    segments = plan_segments(n_frames, n_segments, preroll_frames)
    mine = shard_segments(segments, rank, world)
    local = encode_segments(pcm, mine, encode_batch) if mine else {}
    allp = gather_payloads(local, len(segments), frame_bytes, max(s.n_frames for s in segments), device)
    if allp is None:
        return None
    missing = [s.index for s in segments if s.index not in allp]
    if missing:
        raise RuntimeError("segments missing after the gather: %s" % missing)
    return stitch(segments, allp)


def frame_identity(a, b, frame_bytes):
    """fraction of frames of byte stream `a` that are byte-identical in `b` (same position), and the list of
    differing frame indices"""
    n = (max(len(a), len(b)) + frame_bytes - 1) // frame_bytes
    diff = [k for k in range(n) if a[k * frame_bytes:(k + 1) * frame_bytes] != b[k * frame_bytes:(k + 1) * frame_bytes]]
    return 1.0 - len(diff) / max(n, 1), diff
