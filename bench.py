#!/usr/bin/env python3
"""bench.py — x realtime (audio-s/s) of the Layer III hot path (psy + filterbank + MDCT + rate loop + bitstream) on B200.

Default workload = BASELINE.json configs[3] AS WRITTEN: a batch of 10 000 x 10 s 44.1 kHz stereo clips at 128 kbps,
sharded over the N ranks (10 000 / N clips per GPU: "scaling": "strong").  The clips are heterogeneous synthetic
programme (synth.hetero_batch: the config-1 recipe at levels spread over 30 dB, sums of partials, transients on near
silence, loud noise, digital silence followed by music, a -40 dB clip, an amplitude-modulated tone); clip i is a pure
function of i, so a rank's shard is the same whatever N is.  Each clip is 441 000 samples = 383 frames, the last one
zero-filled as the reference does (encode.c:162-166).  One "step" = one pass of the whole hot path over the rank's shard.

  value : whole-job throughput with the PCM already resident in HBM (mp3gpu_encode_frames_mp3_dev: psy, filterbank,
          MDCT, rate loop + reservoir, device bitstream formatter; MP3 bytes land in a device buffer); profiling off
  e2e   : the same through the reference-facing C ABI with HOST buffers (mp3gpu_encode_frames_mp3): pinned host PCM
          -> H2D -> kernels -> D2H of the finished MP3 byte streams, all inside the timed region
  roofline : the fused polyphase+MDCT front-end kernel: algorithmic bytes / its CUDA-event time, measured in a
          SEPARATE profiled pass (events around every launch), against MEASURED_PEAKS.json hbm_gbs
  parity : after the timed region, k clips of the batch go through the unmodified reference CLI (oracle/_ref/encode) on
          the host: fraction of byte-identical frames, identical streams, decoded SNR (oracle/mp3dec.py) — checker leg
  cpu_baseline : the unmodified reference CLI, one process per host core, one clip of the same batch each

--config 1|2|3 : BASELINE configs[0..2] as batches of 30 s clips; --config 5 : configs[4], ONE 1-hour stream cut into
segments over the ranks with the NCCL gather of the byte streams inside the timed region.
`--impl reference` times the reference's own CPU implementation (all host cores) on the same config.
"""
import argparse
import json
import os
import struct
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic bytes per granule-channel.  front: FP64 xr (5764) or FP32 xr (3460)
BYTES_PER_GC = {"front_polyphase_mdct": 5764, "psy_front": 1152 + 2864, "psy_scan": 1280 + 472,
                "rate_loop": 4608 + 472 + 1152 + 80 + 40, "bitstream": 1152 + 80 + 40 + 417 // 4}
# DRAM bytes per granule-channel of the front-end kernel from the ncu --set full capture (dram__bytes_read + write)
FRONT_TRAFFIC_PER_GC = (0.654215936e9 + 2.239402e9) / 497280
FRONT_TRAFFIC_SOURCE = "profiles/r02_final_ncu_metrics.txt"

CONFIGS = {
    1: dict(fs=44100, n_ch=2, kbps=128, seconds=30.0, clips=2048, cls=0,
            name="batch of %d x 30 s synthetic 44.1 kHz stereo clips at 128 kbps, config-1 recipe (BASELINE configs[0] as a batch)"),
    2: dict(fs=32000, n_ch=1, kbps=64, seconds=30.0, clips=4096, cls=3,
            name="batch of %d x 30 s 32 kHz mono 64 kbps transient-heavy clips (BASELINE configs[1] as a batch)"),
    3: dict(fs=48000, n_ch=2, kbps=320, seconds=30.0, clips=2048, cls=2,
            name="batch of %d x 30 s 48 kHz stereo 320 kbps music-like clips (BASELINE configs[2] as a batch; plain stereo: the "
                 "reference refuses joint stereo for Layer III)"),
    4: dict(fs=44100, n_ch=2, kbps=128, seconds=10.0, clips=10000, cls=None,
            name="batch of %d x 10 s synthetic 44.1 kHz stereo clips at 128 kbps, heterogeneous content, sharded over the GPUs "
                 "(BASELINE configs[3])"),
    5: dict(fs=44100, n_ch=2, kbps=128, seconds=3600.0, clips=1, cls=7,
            name="single %d-s 44.1 kHz stereo stream at 128 kbps segmented at frame boundaries across the GPUs, byte streams "
                 "gathered on rank 0 (BASELINE configs[4])"),
}


def bind_to_gpu_numa_node(torch, index):
    """Run this rank (and first-touch its pinned buffers) on the CPUs of the NUMA node its GPU hangs off: with one rank per
    GPU the uploads of 8 ranks otherwise cross the socket interconnect at random.  Returns the node or None (no-op)."""
    try:
        p = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


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
    return float(t.item()), float(u.item())


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md recipe)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()
        self.proc = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                if self._halt.is_set():
                    break
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self._halt.set()
        if self.proc:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# host-side helpers of the checker legs (reference CLI, decoder): never inside a timed GPU region
# ---------------------------------------------------------------------------------------------------------------------
def wav_bytes(pcm, fs):
    """44-byte-header WAV as the reference sniffs it (musicin.c:352-368); pcm int16 [n_ch][n]"""
    n_ch = pcm.shape[0]
    inter = np.ascontiguousarray(pcm.T).reshape(-1).astype("<i2")
    return b"RIFF" + struct.pack("<I", 36 + inter.nbytes) + b"WAVEfmt " + \
        struct.pack("<IHHIIHH", 16, 1, n_ch, fs, fs * 2 * n_ch, 2 * n_ch, 16) + b"data" + struct.pack("<I", inter.nbytes) + inter.tobytes()


def cli_flags(cfg):
    return (["-m", "m"] if cfg["n_ch"] == 1 else ["-m", "s"]) + ["-s", {32000: "32", 44100: "44.1", 48000: "48"}[cfg["fs"]],
                                                                  "-b", str(cfg["kbps"])]


def reference_encode_many(cfg, pcms, cores=None):
    """run the reference encoder on each PCM array (int16 [n_ch][n]) -> (list of byte streams as the CLI writes them minus
    its spurious last byte, wall seconds, kind).  oracle/_ref/encode when it was built here, else the oracle port."""
    cores = cores or os.cpu_count() or 1
    enc = os.path.join(ROOT, "oracle", "_ref", "encode")
    if os.path.exists(enc):
        tmp = tempfile.mkdtemp(prefix="mp3ref_")
        for i, p in enumerate(pcms):
            with open(os.path.join(tmp, "c%d.wav" % i), "wb") as f:
                f.write(wav_bytes(p, cfg["fs"]))
        t0 = time.perf_counter()
        running, nxt, outs = [], 0, [None] * len(pcms)
        while nxt < len(pcms) or running:
            while nxt < len(pcms) and len(running) < cores:
                cmd = [enc] + cli_flags(cfg) + [os.path.join(tmp, "c%d.wav" % nxt), os.path.join(tmp, "o%d.mp3" % nxt)]
                running.append((nxt, subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)))
                nxt += 1
            i, p = running.pop(0)
            p.wait()
        wall = time.perf_counter() - t0
        for i in range(len(pcms)):
            outs[i] = open(os.path.join(tmp, "o%d.mp3" % i), "rb").read()[:-1]      # close_bit_stream_w's extra byte, common.c:968-974
        subprocess.run(["rm", "-rf", tmp])
        return outs, wall, "reference"
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import multiprocessing as mp
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(min(cores, len(pcms))) as pool:
        outs = pool.starmap(_oracle_job, [(p, cfg["fs"], cfg["kbps"]) for p in pcms])
    return outs, time.perf_counter() - t0, "port"


def _oracle_job(pcm, fs, kbps):
    import oracle
    data, _ = oracle.format_stream(oracle.encode_stream(pcm, fs, kbps), pcm.shape[0], fs, kbps)
    return data[:-1]


def _decode_job(args):
    data, pcm = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mp3dec
    _, dec, ok = mp3dec.decode(data)
    return dec, bool(ok.all()), mp3dec.snr_vs_original(pcm, dec)


def parity_report(cfg, pcms, ours, frame_bytes, n_snr=4):
    """ours[i] (bytes) against the reference bitstream of pcms[i]: identical-frame fraction, identical streams, decoded SNR"""
    refs, wall, kind = reference_encode_many(cfg, pcms)
    frames = same = streams_same = 0
    first_bad = None
    for i, (a, b) in enumerate(zip(ours, refs)):
        n = (max(len(a), len(b)) + frame_bytes - 1) // frame_bytes
        ok = sum(1 for k in range(n) if a[k * frame_bytes:(k + 1) * frame_bytes] == b[k * frame_bytes:(k + 1) * frame_bytes])
        frames += n
        same += ok
        streams_same += int(a == b)
        if ok != n and first_bad is None:
            first_bad = i
    rep = {"reference": kind, "clips": len(pcms), "frames": frames, "identical_frame_fraction": same / max(frames, 1),
           "identical_streams": streams_same}
    sel = list(range(min(n_snr, len(pcms))))
    if first_bad is not None and first_bad not in sel:
        sel.append(first_bad)
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import multiprocessing as mp
        import mp3dec
        jobs = [(ours[i], pcms[i]) for i in sel] + [(refs[i], pcms[i]) for i in sel if refs[i] != ours[i]]
        with mp.get_context("fork").Pool(min(len(jobs), os.cpu_count() or 1)) as pool:
            res = pool.map(_decode_job, jobs)
        mine = res[:len(sel)]
        theirs, k = [], len(sel)
        for j, i in enumerate(sel):
            if refs[i] != ours[i]:
                theirs.append(res[k]); k += 1
            else:
                theirs.append(mine[j])
        vs = [mp3dec.snr_db(t[0], m[0]) for m, t in zip(mine, theirs)]
        finite = [v for v in vs if np.isfinite(v)]
        rep.update(decoded_clips=len(sel), decodable=bool(all(m[1] for m in mine)),
                   decoded_snr_db=float(np.mean([m[2] for m in mine])), decoded_snr_reference_db=float(np.mean([t[2] for t in theirs])),
                   decoded_snr_vs_reference_decode_db=(float(min(finite)) if finite else None),
                   identical_decodes=len(vs) - len(finite))
    except Exception as e:     # the SNR part needs liboracle.so; the byte comparison above does not
        rep["decode_error"] = str(e)[:200]
    return rep


def make_clips_cpu(mod, cfg, indices, n_samples):
    import torch
    out = []
    for i in indices:
        out.append(mod.synth.hetero_batch(torch, int(i), 1, n_samples, cfg["fs"], cfg["n_ch"], "cpu", force_class=cfg["cls"])[0].numpy())
    return out


def run_reference_cpu(mod, cfg, n_samples, cores=None):
    """the unmodified reference CLI, one process per host core, one clip of the bench batch each"""
    cores = cores or os.cpu_count() or 1
    pcms = make_clips_cpu(mod, cfg, range(cores), n_samples)
    _, wall, kind = reference_encode_many(cfg, pcms, cores)
    audio = cores * n_samples / cfg["fs"]
    sample = "%d processes x 1 clip of %.1f s (%d Hz, %d ch, %d kbps; clips 0..%d of the bench batch; whole encoder incl. bitstream " \
             "formatting, process start and table initialisation per clip as the reference runs)" % (
                 cores, n_samples / cfg["fs"], cfg["fs"], cfg["n_ch"], cfg["kbps"], cores - 1)
    return audio / wall, cores, kind, sample, wall


# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=[1, 2, 3, 4, 5], help="BASELINE.json configs[config - 1]; default 4 = configs[3]")
    ap.add_argument("--clips", type=int, default=int(os.environ.get("MP3GPU_BENCH_CLIPS", 0)), help="total clips of the batch (all ranks)")
    ap.add_argument("--seconds", type=float, default=0.0, help="clip (or stream) length; default: the config's")
    ap.add_argument("--chunk-frames", type=int, default=0, help="frames per stream per library call (default: 30 = whole 15-granule tiles of the filterbank kernel for full batches, 192 or 64 below one wave)")
    ap.add_argument("--parity-clips", type=int, default=32, help="clips checked against the reference CLI after the timed region")
    ap.add_argument("--segment-frames", type=int, default=116, help="config 5: frames per segment")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the A/B pass over the front-end variants")
    ap.add_argument("--front", default=os.environ.get("MP3GPU_FRONT", ""), help="front-end variant (see mp3gpu.h); default: the library's")
    ap.add_argument("--pipeline", default=os.environ.get("MP3GPU_PIPELINE", "overlap"), choices=["serial", "overlap"],
                    help="overlap: the front end of chunk i+1 runs beside the rate loop of chunk i (mp3gpu_set_pipeline)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    cfg = dict(CONFIGS[args.config])
    if args.seconds:
        cfg["seconds"] = args.seconds
    if args.clips and args.config != 5:
        cfg["clips"] = args.clips
    FS, NCH, KBPS = cfg["fs"], cfg["n_ch"], cfg["kbps"]
    n_samples = int(round(cfg["seconds"] * FS))
    n_frames = (n_samples + 1151) // 1152              # the last frame is zero-filled (encode.c:162-166): 383 for 10 s at 44.1 kHz
    workload = cfg["name"] % (cfg["clips"] if args.config != 5 else int(cfg["seconds"]))
    config_common = {"workload": workload, "baseline_config": "configs[%d]" % (args.config - 1), "sfreq_hz": FS, "channels": NCH,
                     "bitrate_kbps": KBPS, "clip_seconds": cfg["seconds"], "frames_per_clip": n_frames}

    import mp3gpu_pkg
    mod = mp3gpu_pkg.load()

    if args.impl == "reference":
        if rank != 0:
            return
        steps, vals, walls = max(1, args.steps), [], []
        # bounded sample of the workload: one clip per host core and step (config 5: a 20 s piece of the stream per core)
        ns = n_samples if args.config != 5 else 20 * FS
        for i in range(args.warmup + steps):
            xrt, cores, kind, sample, wall = run_reference_cpu(mod, cfg, ns)
            if i >= args.warmup:
                vals.append(xrt)
                walls.append(wall)
        v = float(np.mean(vals))
        print(json.dumps({"impl": "reference", "metric": "x_realtime", "value": v, "unit": "audio-s/s", "n_gpus": args.gpus,
                          "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(walls)), "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_common,
                          "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": cores, "kind": kind, "sample": sample},
                          "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa_node = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    if args.config == 5:
        import bench_stream
        return bench_stream.run(args, cfg, config_common, mod, rank, world, local_rank, device)

    lo, hi = shard_range(cfg["clips"], rank, world)
    S = hi - lo
    # frames per library call: 32 for a batch that fills the device; a shard below one wave of streams gets longer calls so
    # that the library can cut them into speculative rate-loop segments (mp3gpu_set_rate_loop_segments)
    wave = mod.host.stream_wave(local_rank)
    # full batches: whole filterbank tiles per call.  Below one wave: when the rate loop can cut a stream into segments (needs
    # 2 S <= wave) the calls are as long as the clip (up to 400 frames: the segments get longer, and the library pipelines
    # the upload of a long call with its front end over groups of streams); else medium calls
    chunk = args.chunk_frames or (30 if S >= wave else min(n_frames, 400) if 2 * S <= wave else 64)
    F = min(chunk, n_frames)
    chunks = []
    f0 = 0
    while f0 < n_frames:
        chunks.append((f0, min(F, n_frames - f0)))
        f0 += F
    enc = mod.Encoder(FS, NCH, KBPS, max_streams=S, max_frames=F, device=local_rank)
    if args.front:
        enc.set_front_variant(args.front)
    enc.set_pipeline(args.pipeline == "overlap")
    # ---- synthetic PCM: generated on the device, kept (a) on the device chunk-major for `value`,
    #      (b) in pinned host memory chunk-major for `e2e`
    pcm_all = torch.zeros((S, NCH, n_frames * 1152), dtype=torch.int16, device=device)
    mod.synth.hetero_batch(torch, lo, S, n_samples, FS, NCH, device, out=pcm_all[:, :, :n_samples], force_class=cfg["cls"])
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
    lengths_last = [None]

    def step_dev():
        enc.reset(stream=sptr)
        for i in range(len(chunks)):
            enc.encode_frames_mp3_dev(dev_chunks[i], mp3_dev, stream=sptr)
        return enc.flush_mp3(mp3_dev, S, stream=sptr)

    # e2e: TWO contexts on two streams take the steps in turn (what a host that encodes batch after batch does): the upload
    # and front end of step i + 1 are enqueued before the flush of step i blocks the host, so a step's first upload is not
    # exposed and the device never idles at a step boundary.  Every copy of every step is inside the timed region.
    enc_b = mod.Encoder(FS, NCH, KBPS, max_streams=S, max_frames=F, device=local_rank)
    if args.front:
        enc_b.set_front_variant(args.front)
    enc_b.set_pipeline(args.pipeline == "overlap")
    stream_b = torch.cuda.Stream(device)
    mp3_host_b = torch.zeros((S, mp3_bytes), dtype=torch.uint8, pin_memory=True)
    host_ctx = [(enc, sptr, mp3_host), (enc_b, stream_b.cuda_stream, mp3_host_b)]
    last_host = [0]

    def host_enqueue(i):
        e, sp, out = host_ctx[i % 2]
        e.reset(stream=sp)
        for ch in host_chunks:
            e.encode_frames_mp3(ch.numpy(), out.numpy(), stream=sp)

    def host_finish(i):
        e, sp, out = host_ctx[i % 2]
        lengths_last[0] = e.flush_mp3(out.numpy(), S, stream=sp)
        last_host[0] = i % 2

    def steps_host(k):
        for i in range(k):
            host_enqueue(i)
            if i > 0:
                host_finish(i - 1)
        host_finish(k - 1)

    def timed_host(k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        stream_b.wait_event(e0)
        steps_host(k)
        stream.wait_stream(stream_b)
        e1.record(stream)
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        return e0.elapsed_time(e1) * 1e-3

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

    warmup = max(3, args.warmup)
    for _ in range(warmup):
        step_dev()
    torch.cuda.synchronize(device)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    l0 = enc.kernel_launches
    t_dev = timed(step_dev, args.steps)                  # `value`: profiling off
    launches = enc.kernel_launches - l0
    enc.set_host_delivery(True)     # the D2H of a chunk's bytes overlaps the next chunk's kernels; flush_mp3 joins (mp3gpu.h)
    for e_, _, _ in host_ctx:
        e_.set_host_delivery(True)
    steps_host(2)
    torch.cuda.synchronize(device)
    t_host = timed_host(args.steps)
    clocks = sampler.stop() if rank == 0 else None
    # separate profiled pass: CUDA events around every launch of the hot kernels (per-kernel split + the roofline figure);
    # serial pipeline, so that a kernel's events bracket that kernel alone
    enc.set_host_delivery(False)
    enc.set_pipeline(False)
    enc.profile_enable(True)
    enc.profile_collect(reset=True)
    prof_steps = max(1, min(args.steps, 2))
    for _ in range(prof_steps):
        step_dev()
    prof = enc.profile_collect(reset=True)
    enc.profile_enable(False)
    # A/B of the front-end variants (mp3gpu_set_front_variant): one warm-up step + one profiled step each; what changes in
    # the produced byte streams is measured on the device against the default variant's output of the same batch
    variants = {}
    if not args.no_variants and not args.front:
        base_out = mp3_dev.clone()
        fbytes = enc.frame_bytes
        for v in ("fma", "fma_tc", "fp32"):
            enc.set_front_variant(v)
            step_dev()
            enc.profile_enable(True)
            enc.profile_collect(reset=True)
            step_dev()
            pv = enc.profile_collect(reset=True)
            enc.profile_enable(False)
            same = (mp3_dev.view(S, n_frames, fbytes) == base_out.view(S, n_frames, fbytes)).all(dim=2)
            info = enc.front_variant_info()
            ms = pv["front_polyphase_mdct"][0]
            variants[v] = {"front_ms_per_step": ms, "algorithmic_bytes_per_gc": info["bytes_per_gc"],
                           "achieved_gbs": info["bytes_per_gc"] * S * n_frames * 2 * NCH / (ms * 1e-3) / 1e9 if ms > 0 else 0.0,
                           "step_ms": sum(x[0] for x in pv.values()),
                           "identical_frame_fraction_vs_exact": float(same.float().mean().item()),
                           "identical_streams_vs_exact": int(same.all(dim=1).sum().item())}
        enc.set_front_variant("exact")
        del base_out

    audio_rank = S * n_samples / FS * args.steps
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
    front_info = enc.front_variant_info() if hasattr(enc, "front_variant_info") else {"name": "exact", "bytes_per_gc": 5764}
    bpg = dict(BYTES_PER_GC)
    bpg["front_polyphase_mdct"] = front_info["bytes_per_gc"]
    kernels = {}
    tot_ms = sum(v[0] for v in prof.values()) or 1.0
    for name, (ms, n) in prof.items():
        gbs = bpg[name] * gc_per_step * prof_steps / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        kernels[name] = {"ms_per_step": ms / prof_steps, "launches_per_step": n / prof_steps, "share": ms / tot_ms,
                         "algorithmic_bytes_per_gc": bpg[name], "achieved_gbs": gbs, "frac_hbm": gbs / peak}
    fk = kernels["front_polyphase_mdct"]
    gc_per_launch = S * chunks[0][1] * 2 * NCH
    out = {
        "metric": "x_realtime", "value": audio_total / t_dev_max, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": 1e3 * t_dev_max / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(config_common, clips_total=cfg["clips"], clips_this_rank=S, chunk_frames=F, content=mod.synth.HETERO_CLASSES
                       if cfg["cls"] is None else mod.synth.HETERO_CLASSES[cfg["cls"]],
                       precision="fp64 filterbank/MDCT/rate loop, fp32 FFT (as the reference); front-end variant: " + front_info["name"],
                       pipeline=args.pipeline,
                       l2="inputs larger than L2: %.1f GB PCM and %.1f GB of spectra per step and GPU" % (
                           S * n_frames * 1152 * NCH * 2 / 1e9, gc_per_step * 4608 / 1e9)),
        "e2e": {"value": audio_total / t_host_max, "unit": "audio-s/s",
                "h2d_bytes_per_step": int(S * n_frames * 1152 * NCH * 2), "d2h_bytes_per_step": int(S * mp3_bytes + 4 * S),
                "output": "finished MPEG-1 Layer III byte streams (device bitstream formatter), %d bytes per clip" % mp3_bytes,
                "host_pipeline": "two mp3gpu contexts on two streams take the steps in turn: step i + 1 is enqueued before the flush "
                                 "of step i blocks the host; every H2D / D2H copy of every step is inside the timed region",
                "numa_node": numa_node},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_front_tile (fused polyphase filterbank + MDCT + alias reduction), variant " + front_info["name"],
                     "achieved": fk["achieved_gbs"], "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": fk["frac_hbm"],
                     "traffic": FRONT_TRAFFIC_PER_GC * gc_per_launch,
                     "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum per granule-channel (%s) x "
                                       "granule-channels per launch" % FRONT_TRAFFIC_SOURCE,
                     "achieved_per_launch_bytes": front_info["bytes_per_gc"] * gc_per_launch,
                     "measured_in": "separate profiled pass of %d step(s), serial pipeline (CUDA events around every launch)" % prof_steps},
        "kernels": kernels,
        "clocks": clocks,
    }
    seg_stats = enc.rate_loop_segment_stats()
    if seg_stats:
        out["rate_loop_segments"] = {"per_pass": [dict(zip(("frames_encoded", "frames_replayed", "segments_unchanged", "segments_merged_early"), p))
                                                  for p in seg_stats],
                                     "note": "speculative segmentation of the rate loop for a shard below one wave of streams (mp3gpu.h); "
                                             "totals over all steps of this run"}
    if variants:
        for v in variants.values():
            v["frac_hbm"] = v["achieved_gbs"] / peak
        out["front_variants"] = dict(variants, note="same batch, one profiled step each; exact = the default above; identical_* compare the "
                                     "MP3 bytes of all %d clips of this rank with the default variant's" % S)
    if not args.no_parity and args.parity_clips > 0:
        # checker leg, outside every timed region: clips of this batch through the unmodified reference CLI
        k = min(args.parity_clips, S)
        pick = sorted(np.random.default_rng(20261017).choice(S, size=k, replace=False).tolist())
        for c in range(min(8, S)):                       # every content class at least once
            if c not in pick and len(pick) < k + 8:
                pick.append(c)
        pcms = [np.concatenate([h[i].numpy() for h in host_chunks], axis=1)[:, :n_samples] for i in pick]
        mp3_last = host_ctx[last_host[0]][2]
        ours = [mp3_last[i, :int(lengths_last[0][i])].numpy().tobytes() for i in pick]
        out["parity"] = parity_report(cfg, pcms, ours, enc.frame_bytes)
        out["parity"]["clip_indices"] = [lo + i for i in pick]
    if not args.no_cpu_baseline:
        xrt, cores, kind, sample, wall = run_reference_cpu(mod, cfg, n_samples)
        out["cpu_baseline"] = {"value": xrt, "unit": "audio-s/s", "cores": cores, "kind": kind, "sample": sample}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
