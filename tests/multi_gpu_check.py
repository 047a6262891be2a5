#!/usr/bin/env python3
"""Multi-GPU check of the single-long-stream path (BASELINE configs[4]) under torchrun, NCCL gather:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29555 tests/multi_gpu_check.py [seconds]
Every rank encodes its share of the segments on its own GPU; rank 0 stitches, compares with the one-GPU whole-stream
encode (identical-frame fraction) and decodes both (test decoder) for the SNR report."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import torch
    import torch.distributed as dist
    import mp3gpu_pkg
    pkg = mp3gpu_pkg.load()
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    fs, br = 44100, 128
    pcm = pkg.synth.config1(seconds, fs, seeds=(6, 7))
    env = 0.6 + 0.4 * np.sin(2 * np.pi * 0.05 * np.arange(pcm.shape[1]) / fs)          # configs[4]: slow amplitude envelope
    pcm = np.clip(np.round(pcm * env), -32768, 32767).astype(np.int16)
    seg = pkg.segment
    FB = pkg.Encoder(fs, 2, br, max_streams=1, max_frames=1, device=local).frame_bytes
    n_seg = 8 * world
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = seg.encode_long_stream(pcm, n_seg, seg.gpu_batch_encoder(pkg, fs, 2, br, device=local, chunk_frames=32), FB,
                                 rank=rank, world=world, device=dev if world > 1 else None)
    torch.cuda.synchronize()
    t_seg = time.perf_counter() - t0
    if rank == 0:
        enc = pkg.Encoder(fs, 2, br, max_streams=1, max_frames=32, device=local)
        t0 = time.perf_counter()
        whole = enc.encode_streams(pcm[None])[0]
        t_whole = time.perf_counter() - t0
        frac, diff = seg.frame_identity(whole, out, FB)
        res = {"check": "single stream segmented over GPUs", "n_gpus": world, "segments": n_seg, "audio_s": seconds,
               "frames": (pcm.shape[1] + 1151) // 1152, "identical_frame_fraction": frac, "bytes_whole": len(whole), "bytes_segmented": len(out),
               "wall_s_segmented": t_seg, "wall_s_one_stream_one_gpu": t_whole, "x_realtime_segmented": seconds / t_seg,
               "x_realtime_one_stream": seconds / t_whole}
        if seconds <= 90:
            import mp3dec
            _, dw, okw = mp3dec.decode(whole)
            _, dc, okc = mp3dec.decode(out)
            res.update(decodable=bool(okw.all() and okc.all()), snr_whole_db=mp3dec.snr_vs_original(pcm, dw),
                       snr_segmented_db=mp3dec.snr_vs_original(pcm, dc), snr_segmented_vs_whole_db=mp3dec.snr_db(dw, dc))
        print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
