"""CPU: host-side logic — frame geometry, synthetic generators, stream sharding across ranks (gloo, world 2)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frame_geometry_matches_reference_formula():
    import ref_harness
    # musicin.c:562-572,729-746: 44.1 kHz stereo 128 kbps -> 417 slots, 3336 bits, mean_bits 1524
    assert ref_harness.frame_geometry(44100, 2, 128) == (417, 3336, 1524)
    assert ref_harness.frame_geometry(32000, 1, 64) == (288, 2304, 1068)
    assert ref_harness.frame_geometry(48000, 2, 320) == (960, 7680, 3696)


def test_synth_deterministic(pkg):
    a = pkg.synth.config1(0.2)
    b = pkg.synth.config1(0.2)
    assert a.dtype == np.int16 and a.shape == (2, 8820) and np.array_equal(a, b)
    assert pkg.synth.config2(0.2).shape[0] == 1
    assert np.abs(pkg.synth.full_scale_tone(0.1)).max() >= 32766


def test_shard_plan():
    sys.path.insert(0, ROOT)
    import bench
    for n, w in ((10000, 8), (10, 4), (7, 2), (1, 1)):
        parts = [bench.shard_range(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_reduction():
    """the N>1 bench plumbing (barrier, max-over-ranks timing, unit count all-reduce) on CPU with gloo"""
    code = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
import bench
dist.init_process_group('gloo')
r, w = dist.get_rank(), dist.get_world_size()
lo, hi = bench.shard_range(101, r, w)
t, units = bench.reduce_timing(0.5 + r, hi - lo, None)
assert abs(t - 1.5) < 1e-9 and units == 101, (t, units)
dist.barrier(); dist.destroy_process_group()
print('ok', r)
""" % ROOT
    import tempfile
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
        f.write(code)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", f.name], capture_output=True, text=True, env=env, timeout=300)
    os.unlink(f.name)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.count("ok") == 2


def test_c_host_builds_and_has_no_cpu_fallback(tmp_path):
    """examples/mp3gpu_encode.c compiles with plain gcc against include/mp3gpu.h (no CUDA headers) and, on a machine without
    a CUDA device, stops with the library's error — the hot path never falls back to the CPU"""
    import torch
    exe = os.path.join(ROOT, "examples", "mp3gpu_encode")
    subprocess.run(["gcc", "-O2", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "mp3gpu_encode.c"), "-o", exe, "-L" + os.path.join(ROOT, "mp3-enc-bsd_b200"),
                    "-lmp3gpu", "-Wl,-rpath,$ORIGIN/../mp3-enc-bsd_b200"], check=True)
    wav = tmp_path / "a.wav"
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import write_wav
    write_wav(str(wav), np.zeros((2, 2304), np.int16), 44100)
    p = subprocess.run([exe, str(tmp_path), str(wav)], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert p.returncode == 0, p.stderr
        assert (tmp_path / "a.mp3").stat().st_size == 2 * 417 + 1 - 0 or (tmp_path / "a.mp3").exists()
    else:
        assert p.returncode == 1 and "mp3gpu_create" in p.stderr, (p.returncode, p.stderr)
        assert not (tmp_path / "a.mp3").exists()
