#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libref.so driven by
oracle/ref_harness.py).  Run in the build container only (needs /root/reference to have been compiled
by `make -C oracle ref`).  The fixtures pin the oracle and the CUDA path on machines without the
reference: per-frame side info, scalefactors, ix, pe, ratios, block types, MP3 bytes for ~1 s of each
SURVEY §8d config, plus subband samples / xr of the first frames.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mp3gpu_pkg  # noqa: E402
import ref_harness  # noqa: E402

synth = mp3gpu_pkg.load().synth

CASES = {
    "cfg1_44k_stereo_128": (lambda: synth.config1(1.2), 44100, 128),
    "cfg2_32k_mono_64": (lambda: synth.config2(1.6), 32000, 64),
    "cfg3_48k_stereo_320": (lambda: synth.config3(1.0), 48000, 320),
    "loud_44k_stereo_128": (lambda: synth.loud_sweep(0.8), 44100, 128),
    "scfsi_44k_stereo_128": (lambda: synth.full_scale_tone(0.6), 44100, 128),
}


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, (gen, sfreq, br) in CASES.items():
        pcm = gen()
        r = ref_harness.run_ref_stream(pcm, sfreq, br)
        keep = {k: r[k] for k in ("pe", "ratio_l", "ratio_s", "block_type", "gi", "scalefac_l", "scalefac_s", "scfsi",
                                   "resv_drain", "main_data_begin_next", "mp3")}
        keep["ix"] = r["ix"].astype(np.int16)
        keep["sb_head"] = r["sb"][:6]
        keep["xr_head"] = r["xr"][:6]
        keep["xr_absmax"] = np.abs(r["xr"]).max(axis=-1)
        keep["xr_sum"] = r["xr"].sum(axis=-1)
        keep["pcm"] = pcm
        keep["sfreq"] = np.int32(sfreq)
        keep["bitrate"] = np.int32(br)
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **keep)
        print(name, "frames", len(r["pe"]), "short gc", int((r["block_type"] == 2).sum()), "scfsi", int(r["scfsi"].sum()),
              os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
