"""GPU: 30-second streams against the UNMODIFIED reference CLI, stored as per-frame hashes (tests/golden/long_hashes.json,
tools/make_long_golden.py): the three configurations of BASELINE.json plus loud noise and stereo transients — 1149 / 834 /
1250 frames each, i.e. the reservoir recurrence, scfsi, block switching and table selection over thousands of granules.
Every frame must be byte-identical (exact front end), and the whole-stream digest must match."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle
from util import ROOT

pytestmark = pytest.mark.gpu
G = json.load(open(os.path.join(ROOT, "tests", "golden", "long_hashes.json")))


@pytest.mark.parametrize("name", sorted(G))
def test_30s_stream_every_frame_identical_to_reference_cli(pkg, name):
    g = G[name]
    pcm = pkg.synth.exact_clip(g["kind"], g["seconds"], g["sfreq"], g["n_ch"], g["seed"])
    assert hashlib.sha256(pcm.tobytes()).hexdigest() == g["pcm_sha256"], "exact_clip must give the same samples on every machine"
    enc = pkg.Encoder(g["sfreq"], g["n_ch"], g["bitrate"], max_streams=1, max_frames=64)
    got = enc.encode_streams(pcm[None], chunk_frames=64)[0]
    fb = enc.frame_bytes
    mine = [hashlib.sha256(got[i:i + fb]).hexdigest()[:8] for i in range(0, len(got), fb)]
    ref = [g["frames"][i:i + 8] for i in range(0, len(g["frames"]), 8)]
    same = sum(a == b for a, b in zip(mine, ref))
    print(f"{name}: {same}/{len(ref)} frames byte-identical to the reference CLI ({len(got)} bytes)")
    assert len(got) == g["bytes"] and same == len(ref) == len(mine)
    assert hashlib.sha256(got).hexdigest() == g["sha256"]


def test_long_cases_exercise_short_blocks_and_reservoir():
    """the fixtures are only worth their name if they leave the easy path: the oracle's view of two of them"""
    import mp3gpu_pkg
    synth = mp3gpu_pkg.load().synth
    g = G["cfg2_transient_30s"]
    o = oracle.encode_stream(synth.exact_clip(g["kind"], 6.0, g["sfreq"], g["n_ch"], g["seed"]), g["sfreq"], g["bitrate"])
    bt = o["block_type"][:, :, :g["n_ch"]].reshape(-1)
    assert (bt == 2).sum() >= 20 and (bt == 1).sum() >= 10 and (bt == 3).sum() >= 10, np.bincount(bt, minlength=4)
    assert o["resv_size"].max() > 0
