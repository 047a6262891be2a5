#!/usr/bin/env python3
"""bench.py — x realtime (audio-s/s) of the Layer III hot path (psy + filterbank + MDCT + rate loop) on B200.

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): a batch of independent
10 s 44.1 kHz stereo clips at 128 kbps, synthetic (config-1 recipe: 440 Hz tone + FM tone + noise, distinct
seed per clip).  One "step" = one pass of the whole hot path over the rank's batch.  Streams are independent,
so ranks just take their own batch — no data-path collective, "scaling": "weak" (per-GPU batch fixed).

  value : whole-job throughput with the PCM already resident in HBM (mp3gpu_encode_frames_mp3_dev: psy, filterbank,
          MDCT, rate loop + reservoir, and the device bitstream formatter; MP3 bytes land in a device buffer)
  e2e   : the same through the reference-facing C ABI with HOST buffers (mp3gpu_encode_frames_mp3): pinned
          host PCM -> H2D -> kernels -> D2H of the finished MP3 byte streams, all inside the timed region
  roofline : fused polyphase+MDCT front-end kernel, algorithmic bytes (5764 B per granule-channel,
          FP64 path) / CUDA-event time of that kernel, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline : the unmodified reference CLI encoder (oracle/_ref/encode), one process per host core

`--impl reference` times the reference's own CPU implementation (all host cores) on the same config.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS, NCH, KBPS = 44100, 2, 128
# psy_front: PCM in, PsyMid out (2864 B: partition energies, spread energy, unpredictability, short-transform energies);
# psy_scan: the 1280-byte hot part of PsyMid in, PsyOut (472 B) out
BYTES_PER_GC = {"front_polyphase_mdct": 5764, "psy_front": 1152 + 2864, "psy_scan": 1280 + 472,
                "rate_loop": 4608 + 472 + 1152 + 80 + 40, "bitstream": 1152 + 80 + 40 + 417 // 4}
# DRAM bytes per granule-channel of k_front_tile from the ncu --set full capture profiles/r01_h_capture.md
# (dram__bytes_read.sum + dram__bytes_write.sum = 0.697168 + 2.390813 GB for one launch of 530 432 gc)
FRONT_TRAFFIC_PER_GC = (0.697168e9 + 2.390813e9) / 530432


def shard_range(n, rank, world):
    """contiguous, balanced partition of n units over `world` ranks"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_timing(seconds, units, device):
    """max time over ranks, sum of units over ranks (no-op without torch.distributed)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return seconds, units
    t = torch.tensor([seconds], dtype=torch.float64, device=device or "cpu")
    u = torch.tensor([units], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), int(round(u.item()))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md recipe)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop = index, [], threading.Event()
        self.proc = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                if self._stop.is_set():
                    break
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self._stop.set()
        if self.proc:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def synth_batch_device(torch, n_streams, n_samples, first_seed, device):
    """config-1 recipe generated on the device (torch RNG; the numpy version is used for parity tests)"""
    out = torch.empty((n_streams, NCH, n_samples), dtype=torch.int16, device=device)
    t = torch.arange(n_samples, dtype=torch.float64, device=device) / FS
    tone = 0.25 * torch.sin(2 * np.pi * 440.0 * t)
    fm = 0.15 * torch.sin(2 * np.pi * 1000.0 * t - (500.0 / 0.3) * torch.cos(2 * np.pi * 0.3 * t))
    base = (tone + fm).to(torch.float32)
    g = torch.Generator(device=device)
    step = 256
    for s0 in range(0, n_streams, step):
        s1 = min(n_streams, s0 + step)
        g.manual_seed(first_seed + s0)
        noise = torch.randn((s1 - s0, NCH, n_samples), generator=g, device=device, dtype=torch.float32)
        # per-clip level and tone detune so that clips differ in more than the noise
        lvl = 0.5 + 0.5 * torch.rand((s1 - s0, 1, 1), generator=g, device=device)
        x = (base[None, None, :] + 0.05 * noise) * lvl
        out[s0:s1] = torch.clamp(torch.round(x * 32767.0), -32768, 32767).to(torch.int16)
    return out


def run_reference_cpu(seconds_per_clip, clips_per_core=1, cores=None):
    """the unmodified reference CLI (oracle/_ref/encode), one process per host core, each encoding
    `clips_per_core` distinct config-1 clips.  Returns (x_realtime, cores, kind, sample description)."""
    import mp3gpu_pkg
    synth = mp3gpu_pkg.load().synth
    cores = cores or os.cpu_count() or 1
    enc = os.path.join(ROOT, "oracle", "_ref", "encode")
    kind = "reference" if os.path.exists(enc) else "port"
    tmp = tempfile.mkdtemp(prefix="mp3ref_")
    import struct
    n_distinct = min(8, cores * clips_per_core)
    wavs = []
    for c in range(n_distinct):
        pcm = synth.config1(seconds_per_clip, FS, (2 * c + 1, 2 * c + 2))
        inter = np.ascontiguousarray(pcm.T).reshape(-1)
        hdr = b"RIFF" + struct.pack("<I", 36 + inter.nbytes) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 2, FS, FS * 4, 4, 16) + \
            b"data" + struct.pack("<I", inter.nbytes)
        path = os.path.join(tmp, "c%d.wav" % c)
        with open(path, "wb") as f:
            f.write(hdr + inter.tobytes())
        wavs.append(path)
    t0 = time.perf_counter()
    if kind == "reference":
        procs = []
        for p in range(cores):
            cmd = " && ".join("%s %s %s/o%d_%d.mp3 >/dev/null 2>&1" % (enc, wavs[(p * clips_per_core + j) % n_distinct], tmp, p, j)
                              for j in range(clips_per_core))
            procs.append(subprocess.Popen(cmd, shell=True))
        for p in procs:
            p.wait()
    else:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import multiprocessing as mp
        import oracle
        oracle.lib()
        pcms = [synth.config1(seconds_per_clip, FS, (2 * c + 1, 2 * c + 2)) for c in range(n_distinct)]
        with mp.get_context("fork").Pool(cores) as pool:
            pool.starmap(_oracle_job, [(pcms[(p * clips_per_core + j) % n_distinct],) for p in range(cores) for j in range(clips_per_core)])
    wall = time.perf_counter() - t0
    audio = cores * clips_per_core * seconds_per_clip
    subprocess.run(["rm", "-rf", tmp])
    sample = "%d processes x %d clip(s) of %.1f s 44.1 kHz stereo 128 kbps (whole encoder incl. bitstream formatting)" % (
        cores, clips_per_core, seconds_per_clip)
    return audio / wall, cores, kind, sample, wall


def _oracle_job(pcm):
    import oracle
    oracle.encode_stream(pcm, FS, KBPS)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=int(os.environ.get("MP3GPU_BENCH_STREAMS", 0)),
                    help="clips per GPU (default: one full wave of the rate loop, mp3gpu_stream_wave(): 4144 on a B200)")
    ap.add_argument("--seconds", type=float, default=10.0, help="clip length")
    ap.add_argument("--chunk-frames", type=int, default=32, help="frames per stream per library call")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    n_frames = int(args.seconds * FS) // 1152          # whole frames per clip (383 for 10 s)
    audio_per_stream = n_frames * 1152 / FS
    wl = lambda n: "batch of %d x %.0f s synthetic 44.1 kHz stereo clips at 128 kbps per GPU (BASELINE configs[3])" % (n, args.seconds)
    workload = wl(args.streams or 4144)

    if args.impl == "reference":
        if rank != 0:
            return
        if not args.streams:
            # same workload name as the GPU arm: its default batch is one full wave of the rate loop on this device
            # (28 warps = streams per SM; computed here so that this arm never loads libmp3gpu.so)
            try:
                import torch
                if torch.cuda.is_available():
                    workload = wl(28 * torch.cuda.get_device_properties(local_rank).multi_processor_count)
            except Exception:
                pass
        steps, vals, walls = max(1, args.steps), [], []
        for i in range(args.warmup + steps):
            xrt, cores, kind, sample, wall = run_reference_cpu(args.seconds)
            if i >= args.warmup:
                vals.append(xrt)
                walls.append(wall)
        v = float(np.mean(vals))
        print(json.dumps({"impl": "reference", "metric": "x_realtime", "value": v, "unit": "audio-s/s", "n_gpus": args.gpus,
                          "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(walls)), "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": workload, "sfreq_hz": FS, "channels": NCH, "bitrate_kbps": KBPS},
                          "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": cores, "kind": kind, "sample": sample},
                          "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return

    import torch
    import torch.distributed as dist
    import mp3gpu_pkg
    mod = mp3gpu_pkg.load()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    if not args.streams:
        args.streams = mod.host.stream_wave(local_rank)
        workload = wl(args.streams)
    S, F = args.streams, min(args.chunk_frames, n_frames)
    chunks = []
    f0 = 0
    while f0 < n_frames:
        chunks.append((f0, min(F, n_frames - f0)))
        f0 += F
    enc = mod.Encoder(FS, NCH, KBPS, max_streams=S, max_frames=F, device=local_rank)
    # ---- synthetic PCM: generated on the device, kept (a) on the device chunk-major for `value`,
    #      (b) in pinned host memory chunk-major for `e2e`
    pcm_all = synth_batch_device(torch, S, n_frames * 1152, 1000 * rank + 1, device)
    dev_chunks = [pcm_all[:, :, a * 1152:(a + n) * 1152].contiguous() for a, n in chunks]
    del pcm_all
    host_chunks = [torch.empty(c.shape, dtype=torch.int16, pin_memory=True) for c in dev_chunks]
    for h, d in zip(host_chunks, dev_chunks):
        h.copy_(d)
    mp3_bytes = n_frames * enc.frame_bytes
    mp3_dev = torch.zeros((S, mp3_bytes), dtype=torch.uint8, device=device)
    mp3_host = torch.zeros((S, mp3_bytes), dtype=torch.uint8, pin_memory=True)
    stream = torch.cuda.current_stream(device)
    sptr = stream.cuda_stream

    def step_dev():
        enc.reset()
        for i, (a, n) in enumerate(chunks):
            enc.encode_frames_mp3_dev(dev_chunks[i], mp3_dev, stream=sptr)
        return enc.flush_mp3(mp3_dev, S, stream=sptr)

    def step_host():
        enc.reset()
        for i, (a, n) in enumerate(chunks):
            enc.encode_frames_mp3(host_chunks[i].numpy(), mp3_host.numpy(), stream=sptr)
        return enc.flush_mp3(mp3_host.numpy(), S, stream=sptr)

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(k):
            fn()
        e1.record(stream)
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        return e0.elapsed_time(e1) * 1e-3

    for _ in range(max(3, args.warmup)):
        step_dev()
    torch.cuda.synchronize(device)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    enc.profile_enable(True)
    enc.profile_collect(reset=True)
    l0 = enc.kernel_launches
    t_dev = timed(step_dev, args.steps)
    launches = enc.kernel_launches - l0
    prof = enc.profile_collect(reset=True)
    enc.profile_enable(False)
    enc.set_host_delivery(True)     # the D2H of a chunk's bytes overlaps the next chunk's kernels; flush_mp3 joins (mp3gpu.h)
    for _ in range(2):
        step_host()
    torch.cuda.synchronize(device)
    t_host = timed(step_host, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    # fingerprint of what the timed steps produced (first 8 clips of the last e2e step): lets A/B runs of library variants
    # (tools/ab_bench.sh) see at once that a "faster" variant encodes something else
    import zlib
    out_crc = zlib.crc32(mp3_host[:min(S, 8)].numpy().tobytes()) & 0xffffffff

    audio_rank = S * audio_per_stream * args.steps
    t_dev_max, audio_total = reduce_timing(t_dev, audio_rank, device)
    t_host_max, _ = reduce_timing(t_host, audio_rank, device)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    gc_per_step = S * n_frames * 2 * NCH
    kernels = {}
    tot_ms = sum(v[0] for v in prof.values()) or 1.0
    for name, (ms, n) in prof.items():
        gbs = BYTES_PER_GC[name] * gc_per_step * args.steps / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        kernels[name] = {"ms_per_step": ms / args.steps, "launches_per_step": n / args.steps, "share": ms / tot_ms,
                         "algorithmic_bytes_per_gc": BYTES_PER_GC[name], "achieved_gbs": gbs, "frac_hbm": gbs / peak}
    fk = kernels["front_polyphase_mdct"]
    out = {
        "metric": "x_realtime", "value": audio_total / t_dev_max, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": 1e3 * t_dev_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "sfreq_hz": FS, "channels": NCH, "bitrate_kbps": KBPS, "streams_per_gpu": S,
                   "frames_per_stream": n_frames, "chunk_frames": F, "precision": "fp64 filterbank/MDCT/rate loop, fp32 FFT (as the reference)",
                   "l2": "inputs larger than L2: %.1f GB PCM and %.1f GB of spectra per step" % (
                       S * n_frames * 1152 * NCH * 2 / 1e9, gc_per_step * 4608 / 1e9)},
        "e2e": {"value": audio_total / t_host_max, "unit": "audio-s/s",
                "h2d_bytes_per_step": int(S * n_frames * 1152 * NCH * 2), "d2h_bytes_per_step": int(S * mp3_bytes + 4 * S),
                "output": "finished MPEG-1 Layer III byte streams (device bitstream formatter), %d bytes per clip" % mp3_bytes,
                "output_crc32_first8": "%08x" % out_crc},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_front (fused polyphase filterbank + MDCT + alias reduction, FP64 exact path)",
                     "achieved": fk["achieved_gbs"], "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": fk["frac_hbm"],
                     "traffic": FRONT_TRAFFIC_PER_GC * S * chunks[0][1] * 2 * NCH,
                     "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum per granule-channel "
                                       "(profiles/r01_h_capture.md) x granule-channels per launch",
                     "achieved_per_launch_bytes": 5764 * S * chunks[0][1] * 2 * NCH},
        "kernels": kernels,
        "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        xrt, cores, kind, sample, wall = run_reference_cpu(args.seconds)
        out["cpu_baseline"] = {"value": xrt, "unit": "audio-s/s", "cores": cores, "kind": kind, "sample": sample}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
