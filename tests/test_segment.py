"""CPU: the single-long-stream path (BASELINE configs[4]) — segment planning, sharding over ranks, batching by shape,
the one gather (gloo, world 2) and stitching — with a stand-in batch encoder whose frame bytes depend only on that
frame's PCM, so the stitched result must equal the one-shot result.  The real encoder runs in tests/test_gpu_segment.py."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FB = 24


def fake_batch_encoder(calls=None):
    def encode_batch(batch, preroll):
        S, n_ch, n = batch.shape
        F = n // 1152 - preroll
        out = np.zeros((S, F * FB), np.uint8)
        for s in range(S):
            for f in range(F):
                fr = batch[s, :, (preroll + f) * 1152:(preroll + f + 1) * 1152].astype(np.int64)
                out[s, f * FB:(f + 1) * FB] = [(int(fr.sum()) + 31 * k + int(fr[0, k])) & 255 for k in range(FB)]
        if calls is not None:
            calls.append((S, F, preroll))
        return out, np.full(S, F * FB - 5, np.int64)
    return encode_batch


def test_plan_and_shard(pkg):
    seg = pkg.segment
    for n_frames, n_seg in ((100, 8), (7, 8), (64, 64), (1, 3), (137813, 64)):
        plan = seg.plan_segments(n_frames, n_seg, 2)
        assert plan[0].first_frame == 0 and plan[0].preroll == 0
        assert sum(s.n_frames for s in plan) == n_frames and all(s.n_frames >= 1 for s in plan)
        assert all(a.first_frame + a.n_frames == b.first_frame for a, b in zip(plan, plan[1:]))
        assert all(s.preroll == min(2, s.first_frame) for s in plan)
        assert max(s.n_frames for s in plan) - min(s.n_frames for s in plan) <= 1
        for world in (1, 2, 8):
            parts = [seg.shard_segments(plan, r, world) for r in range(world)]
            assert [s.index for p in parts for s in p] == list(range(len(plan)))
    with pytest.raises(ValueError):
        seg.plan_segments(0, 4)


def test_segment_pcm_preroll_and_zero_padding(pkg):
    seg = pkg.segment
    pcm = np.arange(2 * 5000, dtype=np.int16).reshape(2, 5000)          # 4.34 frames -> 5 frames
    plan = seg.plan_segments(5, 2, 2)                                    # frames [0,3) and [3,5) with 2 pre-roll
    a, b = seg.segment_pcm(pcm, plan[0]), seg.segment_pcm(pcm, plan[1])
    assert a.shape == (2, 3 * 1152) and np.array_equal(a, pcm[:, :3456])
    assert b.shape == (2, 4 * 1152) and np.array_equal(b[:, :5000 - 1152], pcm[:, 1152:]) and not b[:, 5000 - 1152:].any()


def test_stitched_equals_one_shot_single_process(pkg):
    seg = pkg.segment
    rng = np.random.default_rng(5)
    pcm = rng.integers(-3000, 3000, (2, 23 * 1152 - 100), dtype=np.int16)
    calls = []
    whole = seg.encode_long_stream(pcm, 1, fake_batch_encoder(), FB)
    cut = seg.encode_long_stream(pcm, 6, fake_batch_encoder(calls), FB, preroll_frames=2)
    assert whole == cut and len(whole) == 23 * FB - 5
    # 23 frames over 6 segments: sizes 4,4,4,4,4,3; first has no pre-roll -> three batches
    assert sorted(calls) == [(1, 3, 2), (1, 4, 0), (4, 4, 2)]
    frac, diff = seg.frame_identity(whole, cut, FB)
    assert frac == 1.0 and diff == []
    frac, diff = seg.frame_identity(whole, whole[:FB] + b"x" + whole[FB + 1:], FB)
    assert diff == [1]


def test_two_rank_gloo_gather(pkg):
    code = r"""
import os, sys, numpy as np, torch.distributed as dist
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, 'tests'))
import mp3gpu_pkg
from test_segment import fake_batch_encoder, FB
seg = mp3gpu_pkg.load().segment
dist.init_process_group('gloo')
r, w = dist.get_rank(), dist.get_world_size()
pcm = np.random.default_rng(9).integers(-3000, 3000, (1, 37 * 1152), dtype=np.int16)
out = seg.encode_long_stream(pcm, 5, fake_batch_encoder(), FB, rank=r, world=w)
if r == 0:
    assert out == seg.encode_long_stream.__globals__['stitch'](seg.plan_segments(37, 1), seg.encode_segments(pcm, seg.plan_segments(37, 1), fake_batch_encoder()))
else:
    assert out is None
dist.barrier(); dist.destroy_process_group()
print('ok', r)
""" % (ROOT, ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541")
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
        f.write(code)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29541", f.name], capture_output=True, text=True, env=env, timeout=300)
    os.unlink(f.name)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.count("ok") == 2
