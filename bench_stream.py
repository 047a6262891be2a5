"""bench.py --config 5 — BASELINE configs[4]: ONE long stream (default 1 hour, 44.1 kHz stereo, 128 kbps) cut at frame
boundaries into segments that the ranks encode as independent streams; the finished byte streams are gathered on rank 0
(NCCL) INSIDE the timed region and stitched there.

Every stage of the hot path except the bit reservoir has a bounded memory of the signal (segment.py), so a segment that is
fed `preroll` frames of the preceding audio first reproduces spectrum, thresholds and block types of the whole-stream encode;
the reservoir is emptied at the seam (mp3gpu_begin_segment), which is why frames after a seam can differ from the reference
until the recurrence re-converges — the identical-frame fraction against the reference CLI is part of the report.  The first
segment has no pre-roll: its stream is restarted (mp3gpu_reset_streams) so that it begins exactly like the reference's process.

  value : PCM of the rank's part resident in HBM -> segment batch (strided view, torch: plumbing) -> pre-roll call ->
          begin_segment -> encode calls -> flush -> all_gather of the byte streams -> stitched stream in HBM on rank 0
  e2e   : the same from pinned host PCM (H2D inside) to the stitched byte stream in pinned host memory on rank 0 (D2H inside)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def run(args, cfg, config_common, mod, rank, world, local_rank, device):
    import torch
    import torch.distributed as dist
    import bench
    FS, NCH, KBPS = cfg["fs"], cfg["n_ch"], cfg["kbps"]
    n_samples = int(round(cfg["seconds"] * FS))
    n_frames = (n_samples + 1151) // 1152
    L, P = int(args.segment_frames), 8
    n_seg = (n_frames + L - 1) // L
    klo, khi = bench.shard_range(n_seg, rank, world)
    S = khi - klo
    Smax = (n_seg + world - 1) // world
    enc = mod.Encoder(FS, NCH, KBPS, max_streams=max(S, 1), max_frames=max(L, P), device=local_rank)
    if args.front:
        enc.set_front_variant(args.front)
    # serial pipeline: the segment batch is produced on the caller's stream right before each call, and overlap mode reads
    # the PCM on a private stream (mp3gpu.h: the buffer must be complete when the call is made)
    enc.set_pipeline(False)
    FB = enc.frame_bytes
    # the rank's part of the stream: frames [klo*L - P, khi*L), zero-padded at both ends of the stream
    a, b = (klo * L - P) * 1152, khi * L * 1152
    part = torch.zeros((NCH, b - a), dtype=torch.int16, device=device)
    lo_s, hi_s = max(a, 0), min(b, n_samples)
    if hi_s > lo_s:
        # the stream is clip 0 of class "am-tone" (config-1 recipe under a slow amplitude envelope): generate exactly the
        # samples [lo_s, hi_s) — the recipe is a pure function of the absolute sample index
        whole = mod.synth.hetero_stream(torch, lo_s, hi_s - lo_s, FS, NCH, device)
        part[:, lo_s - a:hi_s - a] = whole
        del whole
    part_host = torch.empty(part.shape, dtype=torch.int16, pin_memory=True)
    part_host.copy_(part)
    part_in = torch.empty_like(part)
    frames_of = [min(L, n_frames - k * L) for k in range(klo, khi)]
    seg_len = (P + L) * 1152
    row_bytes = L * FB
    mine = torch.zeros((Smax, row_bytes), dtype=torch.uint8, device=device)
    allrows = torch.zeros((world * Smax, row_bytes), dtype=torch.uint8, device=device)
    out_host = torch.zeros((n_seg * row_bytes,), dtype=torch.uint8, pin_memory=True) if rank == 0 else None
    out_dev = torch.zeros((n_seg * row_bytes,), dtype=torch.uint8, device=device) if rank == 0 else None
    stream = torch.cuda.current_stream(device)
    sptr = stream.cuda_stream
    last_len = [0]

    def step(host):
        src = part
        if host:
            part_in.copy_(part_host, non_blocking=True)         # H2D of the rank's part of the stream
            src = part_in
        if S > 0:
            batch = src.unfold(1, seg_len, L * 1152).permute(1, 0, 2).contiguous()       # [S][n_ch][(P+L)*1152]
            enc.reset(stream=sptr)
            pre = batch[:, :, :P * 1152].contiguous()
            enc.encode_frames_dev(pre, stream=sptr)             # pre-roll: outputs discarded
            enc.begin_segment(stream=sptr)
            if klo == 0:
                enc.reset_streams(0, 1, stream=sptr)            # the stream's very first segment starts from zero state
            enc.set_stream_frames(frames_of, stream=sptr)
            body = batch[:, :, P * 1152:].contiguous()
            enc.encode_frames_mp3_dev(body, mine, stream=sptr)
            lengths = enc.flush_mp3(mine, S, stream=sptr)
            if khi == n_seg:
                last_len[0] = int(lengths[-1])
        if world > 1:
            dist.all_gather_into_tensor(allrows, mine)
        else:
            allrows.copy_(mine)
        if rank == 0:
            # stitch: segment k is row (k - klo_r) of rank r's block; all rows have L*FB bytes, the last is cut by its length
            pos = 0
            for r in range(world):
                rlo, rhi = bench.shard_range(n_seg, r, world)
                n = rhi - rlo
                if n:
                    out_dev[pos * row_bytes:(pos + n) * row_bytes].copy_(allrows[r * Smax:r * Smax + n].reshape(-1))
                    pos += n
            if host:
                out_host.copy_(out_dev, non_blocking=True)

    def timed(host, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(k):
            step(host)
        e1.record(stream)
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        return e0.elapsed_time(e1) * 1e-3

    warmup = max(3, args.warmup)
    for _ in range(warmup):
        step(False)
    torch.cuda.synchronize(device)
    sampler = bench.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    l0 = enc.kernel_launches
    t_dev = timed(False, args.steps)
    launches = enc.kernel_launches - l0
    for _ in range(2):
        step(True)
    t_host = timed(True, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    enc.profile_enable(True)
    enc.profile_collect(reset=True)
    step(False)
    prof = enc.profile_collect(reset=True)
    enc.profile_enable(False)
    # the length of the last segment lives on the last rank
    tl = torch.tensor([last_len[0]], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(tl, op=dist.ReduceOp.MAX)
    audio = n_samples / FS * args.steps
    t_dev_max, _ = bench.reduce_timing(t_dev, 0, device)
    t_host_max, _ = bench.reduce_timing(t_host, 0, device)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    total_len = (n_seg - 1) * row_bytes + int(tl.item())
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    front_info = enc.front_variant_info() if hasattr(enc, "front_variant_info") else {"name": "exact", "bytes_per_gc": 5764}
    bpg = dict(bench.BYTES_PER_GC)
    bpg["front_polyphase_mdct"] = front_info["bytes_per_gc"]
    gc_rank = sum(frames_of) * 2 * NCH + S * P * 2 * NCH
    kernels, tot_ms = {}, sum(v[0] for v in prof.values()) or 1.0
    for name, (ms, n) in prof.items():
        gbs = bpg[name] * gc_rank / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        kernels[name] = {"ms_per_step": ms, "launches_per_step": n, "share": ms / tot_ms, "algorithmic_bytes_per_gc": bpg[name],
                         "achieved_gbs": gbs, "frac_hbm": gbs / peak}
    fk = kernels["front_polyphase_mdct"]
    out = {
        "metric": "x_realtime", "value": audio / t_dev_max, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": 1e3 * t_dev_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": dict(config_common, segments=n_seg, segment_frames=L, preroll_frames=P, segments_this_rank=S,
                       gather="torch.distributed all_gather_into_tensor (NCCL) of %d x %d bytes per rank, inside the timed region" % (Smax, row_bytes),
                       precision="fp64 filterbank/MDCT/rate loop, fp32 FFT (as the reference); front-end variant: " + front_info["name"],
                       l2="the stream (%.0f MB PCM) is smaller than L2 per rank at 8 GPUs; every step re-reads it from HBM after "
                          "%.1f GB of spectra / intermediates went through L2" % (part.numel() * 2 / 1e6, gc_rank * 7500 / 1e9)),
        "e2e": {"value": audio / t_host_max, "unit": "audio-s/s", "h2d_bytes_per_step": int(part.numel() * 2 * world),
                "d2h_bytes_per_step": int(n_seg * row_bytes), "output": "stitched MPEG-1 Layer III byte stream of %d bytes in pinned host memory on rank 0" % total_len},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_front_tile (fused polyphase filterbank + MDCT + alias reduction), variant " + front_info["name"],
                     "achieved": fk["achieved_gbs"], "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": fk["frac_hbm"], "traffic": None,
                     "achieved_per_launch_bytes": front_info["bytes_per_gc"] * gc_rank / max(fk["launches_per_step"], 1),
                     "measured_in": "separate profiled pass (CUDA events around every launch); few streams: launch-latency bound"},
        "kernels": kernels, "clocks": clocks,
    }
    if not args.no_parity:
        # checker leg: the first `pref` seconds of the stream through the unmodified reference CLI.  The encoder is causal, so
        # the reference's file for the prefix equals the prefix of its file for the whole stream up to the last frames.
        pref = min(float(os.environ.get("MP3GPU_BENCH_PREFIX_S", 90.0)), cfg["seconds"])
        npre = int(pref * FS)
        pcm = mod.synth.hetero_stream(torch, 0, npre, FS, NCH, "cpu").numpy()
        refs, wall, kind = bench.reference_encode_many(cfg, [pcm])
        ours = out_host.numpy().tobytes()[:total_len]
        nfr = min(len(refs[0]) // FB - 2, len(ours) // FB)
        same = sum(1 for k in range(nfr) if ours[k * FB:(k + 1) * FB] == refs[0][k * FB:(k + 1) * FB])
        rep = {"reference": kind, "prefix_seconds": pref, "frames": nfr, "seams_in_prefix": max(0, (nfr - 1) // L),
               "identical_frame_fraction": same / max(nfr, 1)}
        try:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import mp3dec
            nd = min(nfr, int(20 * FS / 1152))
            _, d1, ok1 = mp3dec.decode(ours[:nd * FB])
            _, d2, ok2 = mp3dec.decode(refs[0][:nd * FB])
            v = mp3dec.snr_db(d2, d1)
            rep.update(decoded_seconds=nd * 1152 / FS, decodable=bool(ok1.all()), decoded_snr_db=mp3dec.snr_vs_original(pcm, d1),
                       decoded_snr_reference_db=mp3dec.snr_vs_original(pcm, d2),
                       decoded_snr_vs_reference_decode_db=(float(v) if np.isfinite(v) else None))
        except Exception as e:
            rep["decode_error"] = str(e)[:200]
        out["parity"] = rep
    if not args.no_cpu_baseline:
        xrt, cores, kind, sample, wall = bench.run_reference_cpu(mod, cfg, 20 * FS)
        out["cpu_baseline"] = {"value": xrt, "unit": "audio-s/s", "cores": cores, "kind": kind, "sample": sample +
                               " — a single stream is inherently serial for the reference: one core delivers value / cores"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
