"""GPU: breadth — every MPEG-1 sample rate x mono/stereo x low/mid/high bitrate, ragged and tiny inputs, extreme PCM,
and a full-size batch (one whole wave of streams, the bench's shape) checked through a size-independent property.
Everything is compared byte for byte with the oracle's encoder + formatter (itself pinned on the reference, tests/test_oracle.py)."""
import hashlib

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def oracle_bytes(pcm, fs, br):
    data, _ = oracle.format_stream(oracle.encode_stream(pcm, fs, br), pcm.shape[0], fs, br)
    return data[:-1]            # minus the spurious byte of close_bit_stream_w (common.c:968-974)


def mix(pkg, n, fs, n_ch, seed):
    x = pkg.synth.config1(n / fs + 0.01, fs, seeds=(seed, seed + 1))[:n_ch, :n]
    return np.ascontiguousarray(x)


@pytest.mark.parametrize("fs", [32000, 44100, 48000])
@pytest.mark.parametrize("n_ch,br", [(1, 32), (1, 96), (1, 320), (2, 32), (2, 160), (2, 320)])
def test_format_matrix(pkg, fs, n_ch, br):
    pcm = mix(pkg, 10 * 1152, fs, n_ch, 7 + br)
    enc = pkg.Encoder(fs, n_ch, br, max_streams=1, max_frames=4)
    got = enc.encode_streams(pcm[None], chunk_frames=4)[0]
    ref = oracle_bytes(pcm, fs, br)
    assert got == ref, (fs, n_ch, br, len(got), len(ref))


@pytest.mark.parametrize("n", [1, 31, 1151, 1152, 1153, 2 * 1152 + 577])
def test_ragged_and_tiny_inputs(pkg, n):
    """the last frame is zero-filled like get_audio() does (encode.c:162-166); a 1-sample stream is one frame"""
    pcm = mix(pkg, n, 44100, 2, 3)
    enc = pkg.Encoder(44100, 2, 128, max_streams=1, max_frames=2)
    got = enc.encode_streams(pcm[None])[0]
    assert got == oracle_bytes(pcm, 44100, 128), (n, len(got))


def test_extreme_pcm(pkg):
    """full-scale square wave, most negative DC, single impulses, alternating sign at Nyquist: in one batch"""
    n = 6 * 1152
    t = np.arange(n)
    sig = np.zeros((5, 2, n), np.int16)
    sig[0] = np.where((t // 50) % 2 == 0, 32767, -32768)
    sig[1] = -32768
    sig[2, 0, 700] = 32767; sig[2, 1, 3000] = -32768
    sig[3] = np.where(t % 2 == 0, 32767, -32768)
    sig[4, 0] = 32767                                                   # one channel DC, the other silent
    enc = pkg.Encoder(44100, 2, 128, max_streams=5, max_frames=3)
    got = enc.encode_streams(sig, chunk_frames=3)
    for s in range(5):
        assert got[s] == oracle_bytes(sig[s], 44100, 128), s


def test_full_wave_batch_replicas_identical(pkg):
    """BASELINE configs[3] shape at full width: one whole wave of streams (mp3gpu_stream_wave, 3552 on a B200) built
    from 6 distinct clips.  Size-independent property: a stream's bytes depend on nothing but its own PCM, so every
    replica must equal the oracle's bytes for its clip (checksum of checksums over the batch)."""
    S = pkg.host.stream_wave(0)
    F = 20
    clips = [mix(pkg, F * 1152, 44100, 2, 40 + 2 * i) for i in range(6)]
    clips[4][:] = 0
    want = [hashlib.sha256(oracle_bytes(c, 44100, 128)).hexdigest() for c in clips]
    pcm = torch.empty((S, 2, F * 1152), dtype=torch.int16).pin_memory().numpy()
    for s in range(S):
        pcm[s] = clips[s % 6]
    enc = pkg.Encoder(44100, 2, 128, max_streams=S, max_frames=10)
    mp3 = torch.zeros((S, F * enc.frame_bytes), dtype=torch.uint8).pin_memory().numpy()
    for f0 in (0, 10):
        enc.encode_frames_mp3(np.ascontiguousarray(pcm[:, :, f0 * 1152:(f0 + 10) * 1152]), mp3)
    lengths = enc.flush_mp3(mp3, S)
    bad = [s for s in range(S) if hashlib.sha256(mp3[s, :lengths[s]].tobytes()).hexdigest() != want[s % 6]]
    print(f"{S} streams x {F} frames: {S - len(bad)}/{S} byte streams identical to the oracle's")
    assert not bad, bad[:10]
