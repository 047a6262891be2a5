#!/usr/bin/env python3
"""GPU diagnostic: whole pipeline vs oracle + golden, field by field (run under gpurun)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mp3gpu_pkg, oracle
from util import oracle_flat, pad_frames, psy_array
import torch
pkg = mp3gpu_pkg.load()
d = os.path.join(ROOT, "tests", "golden")
for name in sorted(os.listdir(d)):
    g = np.load(os.path.join(d, name))
    pcm, fs, br = g["pcm"], int(g["sfreq"]), int(g["bitrate"])
    n_ch = pcm.shape[0]
    padded, nf = pad_frames(pcm)
    o = oracle_flat(oracle.encode_stream(pcm, fs, br), n_ch)
    enc = pkg.Encoder(fs, n_ch, br, max_streams=1, max_frames=nf)
    out = enc.encode_frames(np.ascontiguousarray(padded[None]))
    ngc = nf * 2 * n_ch
    ref_ix = np.ascontiguousarray(g["ix"][:, :, :n_ch]).reshape(ngc, 576).astype(np.int32)
    ref_gi = np.ascontiguousarray(g["gi"][:, :, :n_ch]).reshape(ngc, 20)
    ix = out["ix"][0].astype(np.int32)
    print(name, "ngc", ngc)
    print("  golden: |ix| ok", (np.abs(ix) == np.abs(ref_ix)).all(axis=1).mean(), "signed ok", (ix == ref_ix).all(axis=1).mean(),
          "gi ok", (out["gi"][0] == ref_gi).all(axis=1).mean())
    print("  oracle: |ix| ok", (np.abs(ix) == o["ix"]).all(axis=1).mean(), "gi ok", (out["gi"][0] == o["gi"]).all(axis=1).mean())
    print("  oracle-vs-golden: |ix|", (o["ix"] == np.abs(ref_ix)).all(axis=1).mean(), "gi", (o["gi"] == ref_gi).all(axis=1).mean())
    bad = np.where(~(out["gi"][0] == o["gi"]).all(axis=1))[0]
    if len(bad):
        b = bad[0]
        print("  first bad gc", b, "gpu gi", out["gi"][0][b].tolist(), "\n    oracle gi", o["gi"][b].tolist())
    # stages
    dev = torch.device("cuda", 0)
    enc2 = pkg.Encoder(fs, n_ch, br, max_streams=1, max_frames=nf)
    pc = torch.from_numpy(padded[None].copy()).to(dev)
    psy = pkg.host.psy_to_numpy(enc2.L3psycho_anal_batch(pc))[0]
    print("  psy: block_type ok", (psy["block_type"] == o["block_type"]).mean(), "pe exact", (psy["pe"] == o["pe"]).mean(),
          "max dpe", np.abs(psy["pe"] - o["pe"]).max(), "ratio_l exact", (psy["ratio_l"] == o["ratio_l"]).mean(),
          "ratio_s exact", (psy["ratio_s"] == o["ratio_s"]).mean())
    sb = enc2.filter_subband_batch(pc).cpu().numpy()[0]
    print("  sb exact", np.array_equal(sb, o["sb"]))
    psy_t = pkg.host.psy_from_numpy(psy_array(o)[None], dev)
    xr = enc2.subband_mdct_batch(pc, psy_t).cpu().numpy()[0]
    print("  xr exact (oracle psy)", np.array_equal(xr, o["xr"]), np.abs(xr - o["xr"]).max())
    r = enc2.iteration_loop_batch(torch.from_numpy(o["xr"][None].copy()).to(dev), psy_t)
    print("  loop (oracle xr/psy): |ix| ok", (np.abs(r["ix"].cpu().numpy()[0].astype(np.int32)) == o["ix"]).all(axis=1).mean(),
          "gi ok", (r["gi"].cpu().numpy()[0] == o["gi"]).all(axis=1).mean())
