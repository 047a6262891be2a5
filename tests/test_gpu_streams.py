"""GPU: independent stream lifetimes in the batched API — streams of different lengths in ONE ctx
(mp3gpu_set_stream_frames), restarted streams (mp3gpu_reset_streams), the stream-ordered reset, the delivery contract of
the pipelined host path, and the persistent work-queue rate loop at batch sizes below, at and far above the number of
warp slots of the device.  Byte streams are compared with the oracle's encoder + formatter (pinned on the reference)."""
import hashlib

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def oracle_bytes(pcm, fs=44100, br=128):
    data, _ = oracle.format_stream(oracle.encode_stream(pcm, fs, br), pcm.shape[0], fs, br)
    return data[:-1]


def clip(pkg, n, seed, n_ch=2, fs=44100):
    return np.ascontiguousarray(pkg.synth.config1(n / fs + 0.01, fs, seeds=(seed, seed + 1))[:n_ch, :n])


@pytest.mark.parametrize("chunk", [3, 8])
def test_mixed_lengths_in_one_ctx(pkg, chunk):
    """7 streams of 1..12 frames (incl. ragged sample counts and an empty one) share one ctx and one lockstep call sequence;
    each must come out exactly as if encoded alone (the reference: one stream of any length per process, musicin.c:585)"""
    lens = [12 * 1152, 5 * 1152 + 17, 1, 7 * 1152, 0, 9 * 1152 - 1, 3 * 1152]
    S, nmax = len(lens), max(lens)
    pcm = np.zeros((S, 2, nmax), np.int16)
    clips = []
    for s, n in enumerate(lens):
        c = clip(pkg, n, 500 + 3 * s) if n else np.zeros((2, 0), np.int16)
        clips.append(c)
        pcm[s, :, :n] = c
    enc = pkg.Encoder(44100, 2, 128, max_streams=S, max_frames=chunk)
    got = enc.encode_streams(pcm, chunk_frames=chunk, n_samples=lens)
    for s, c in enumerate(clips):
        want = oracle_bytes(c) if lens[s] else b""
        assert got[s] == want, (s, lens[s], len(got[s]), len(want))


def test_garbage_beyond_stream_end_is_ignored(pkg):
    """PCM rows beyond a stream's end are never read: fill them with noise, the result must not change"""
    lens = [4 * 1152, 2 * 1152]
    pcm = np.stack([clip(pkg, 4 * 1152, 610), clip(pkg, 4 * 1152, 620)])
    enc = pkg.Encoder(44100, 2, 128, max_streams=2, max_frames=2)
    got = enc.encode_streams(pcm, chunk_frames=2, n_samples=lens)
    assert got[0] == oracle_bytes(pcm[0])
    assert got[1] == oracle_bytes(pcm[1][:, :2 * 1152])


def test_reset_streams_restarts_only_the_given_streams(pkg):
    F = 6
    a, b = clip(pkg, F * 1152, 700), clip(pkg, F * 1152, 710)
    enc = pkg.Encoder(44100, 2, 128, max_streams=2, max_frames=F)
    out0 = enc.encode_frames(np.stack([a, b]))                     # both streams now carry history
    enc.reset_streams(1, 1)
    out1 = enc.encode_frames(np.stack([a, b]))
    # stream 1 was restarted: identical to its first encode; stream 0 continued: differs from a fresh start
    assert np.array_equal(out1["ix"][1], out0["ix"][1]) and np.array_equal(out1["gi"][1], out0["gi"][1])
    assert not np.array_equal(out1["ix"][0], out0["ix"][0])


def test_stream_ordered_reset(pkg):
    S, F = 4, 5
    pcm = np.stack([clip(pkg, F * 1152, 800 + 2 * s) for s in range(S)])
    enc = pkg.Encoder(44100, 2, 128, max_streams=S, max_frames=F)
    st = torch.cuda.Stream()
    a = enc.encode_streams(pcm)
    mp3 = torch.zeros((S, F * enc.frame_bytes), dtype=torch.uint8).pin_memory().numpy()
    enc.reset(stream=st.cuda_stream)
    enc.encode_frames_mp3(pcm, mp3, stream=st.cuda_stream)
    lengths = enc.flush_mp3(mp3, S, stream=st.cuda_stream)
    for s in range(S):
        assert mp3[s, :lengths[s]].tobytes() == a[s]


def test_pipelined_delivery_contract(pkg):
    """mp3gpu.h: with MP3GPU_DELIVER_PIPELINED the bytes of call i have landed once the work of call i+1 has completed on
    its stream — read them then, before any flush"""
    S, F, C = 6, 24, 4
    pcm = np.stack([clip(pkg, F * 1152, 900 + 2 * s) for s in range(S)])
    want = [oracle_bytes(pcm[s]) for s in range(S)]
    enc = pkg.Encoder(44100, 2, 128, max_streams=S, max_frames=C)
    enc.set_host_delivery(True)
    FB, T = enc.frame_bytes, 511 // (enc.frame_bytes - enc.sideinfo_bytes) + 1
    mp3 = torch.zeros((S, F * FB), dtype=torch.uint8).pin_memory().numpy()
    pin = torch.from_numpy(pcm).pin_memory().numpy()
    calls = F // C
    for i in range(calls):
        enc.encode_frames_mp3(np.ascontiguousarray(pin[:, :, i * C * 1152:(i + 1) * C * 1152]), mp3)
        enc.sync()
        if i >= 1:
            # call i-1 delivered absolute frames [(i-1)*C - T, i*C - T): final by construction
            lo, hi = max(0, ((i - 1) * C - T) * FB), max(0, (i * C - T) * FB)
            for s in range(S):
                assert mp3[s, lo:hi].tobytes() == want[s][lo:hi], (i, s)
    lengths = enc.flush_mp3(mp3, S)
    for s in range(S):
        assert mp3[s, :lengths[s]].tobytes() == want[s]


@pytest.mark.parametrize("S", [1, 37, 300, 9000])
def test_work_queue_rate_loop_any_batch_size(pkg, S):
    """the persistent rate loop draws (stream, frame) items from a queue: 1 stream, fewer streams than SMs, fewer than warp
    slots, and more than two waves of them must all give every stream exactly the oracle's bytes for its clip"""
    F = 6
    clips = [clip(pkg, F * 1152, 40 + 2 * i) for i in range(5)]
    clips[3][:] = 0
    want = [hashlib.sha256(oracle_bytes(c)).hexdigest() for c in clips]
    pcm = torch.empty((S, 2, F * 1152), dtype=torch.int16).pin_memory().numpy()
    for s in range(S):
        pcm[s] = clips[s % 5]
    enc = pkg.Encoder(44100, 2, 128, max_streams=S, max_frames=3)
    mp3 = torch.zeros((S, F * enc.frame_bytes), dtype=torch.uint8).pin_memory().numpy()
    for f0 in (0, 3):
        enc.encode_frames_mp3(np.ascontiguousarray(pcm[:, :, f0 * 1152:(f0 + 3) * 1152]), mp3)
    lengths = enc.flush_mp3(mp3, S)
    bad = [s for s in range(S) if hashlib.sha256(mp3[s, :lengths[s]].tobytes()).hexdigest() != want[s % 5]]
    assert not bad, (S, bad[:10])


def test_two_ctxs_on_one_device_interleaved(pkg):
    """entry points bind the ctx's device themselves and restore the caller's (DeviceGuard); two ctxs interleave"""
    a, b = clip(pkg, 4 * 1152, 1000), clip(pkg, 4 * 1152, 1010, n_ch=1)
    e1 = pkg.Encoder(44100, 2, 128, max_streams=1, max_frames=2)
    e2 = pkg.Encoder(44100, 1, 64, max_streams=1, max_frames=2)
    m1 = np.zeros((1, 4 * e1.frame_bytes), np.uint8)
    m2 = np.zeros((1, 4 * e2.frame_bytes), np.uint8)
    for f0 in (0, 2):
        e1.encode_frames_mp3(np.ascontiguousarray(a[None, :, f0 * 1152:(f0 + 2) * 1152]), m1)
        e2.encode_frames_mp3(np.ascontiguousarray(b[None, :, f0 * 1152:(f0 + 2) * 1152]), m2)
    l1, l2 = e1.flush_mp3(m1, 1), e2.flush_mp3(m2, 1)
    assert m1[0, :l1[0]].tobytes() == oracle_bytes(a) and m2[0, :l2[0]].tobytes() == oracle_bytes(b, 44100, 64)
    assert torch.cuda.current_device() == 0


@pytest.mark.parametrize("host", [True, False])
def test_overlap_pipeline_gives_identical_bytes(pkg, host):
    """MP3GPU_PIPELINE_OVERLAP: the front end of call i+1 runs beside the rate loop of call i on a private stream with
    double-buffered spectra; ragged chunks, per-stream lengths, two batches with a stream-ordered reset between them"""
    S, F = 9, 26
    lens = [F * 1152, 20 * 1152 + 5, 3 * 1152, F * 1152, 1, 11 * 1152, F * 1152 - 1, 7 * 1152, 25 * 1152]
    pcm = np.zeros((S, 2, F * 1152), np.int16)
    clips = []
    for s, n in enumerate(lens):
        c = clip(pkg, n, 1200 + 3 * s)
        clips.append(c)
        pcm[s, :, :n] = c
    want = [oracle_bytes(c) for c in clips]
    enc = pkg.Encoder(44100, 2, 128, max_streams=S, max_frames=7)
    enc.set_pipeline(True)
    st = torch.cuda.Stream()
    dev = torch.device("cuda", 0)
    for batch in range(2):
        enc.reset(stream=st.cuda_stream)
        enc.set_stream_frames([(n + 1151) // 1152 for n in lens], stream=st.cuda_stream)
        if host:
            mp3 = torch.zeros((S, F * enc.frame_bytes), dtype=torch.uint8).pin_memory().numpy()
            pin = torch.from_numpy(pcm).pin_memory().numpy()
        else:
            mp3 = torch.zeros((S, F * enc.frame_bytes), dtype=torch.uint8, device=dev)
            pin = torch.from_numpy(pcm).to(dev)
        # overlap mode reads the PCM on a private stream: the chunks must be complete when the call is made and stay alive
        # until the batch is flushed (mp3gpu.h), so they are all prepared up front
        sizes, f0, chunks = [7, 1, 5, 7, 2, 4], 0, []
        for cfr in sizes:
            ch = pin[:, :, f0 * 1152:(f0 + cfr) * 1152]
            chunks.append(np.ascontiguousarray(ch) if host else ch.contiguous())
            f0 += cfr
        assert f0 == F
        torch.cuda.synchronize()
        for ch in chunks:
            if host:
                enc.encode_frames_mp3(ch, mp3, stream=st.cuda_stream)
            else:
                enc.encode_frames_mp3_dev(ch, mp3, stream=st.cuda_stream)
        lengths = enc.flush_mp3(mp3, S, stream=st.cuda_stream)
        out = mp3 if host else mp3.cpu().numpy()
        for s in range(S):
            assert out[s, :lengths[s]].tobytes() == want[s], (batch, s, lengths[s], len(want[s]))
    enc.set_pipeline(False)
    got = enc.encode_streams(pcm, chunk_frames=7, n_samples=lens)
    assert got == want


@pytest.mark.parametrize("S,F,chunk", [(3, 40, 40), (12, 70, 35), (40, 48, 48)])
def test_speculative_rate_loop_segments_are_bit_identical(pkg, S, F, chunk):
    """under-filled batches cut every call into up to 8 concurrently encoded segments per stream with a guessed reservoir
    and re-encode what started wrong (k_rate_loop): every byte, and the state carried into the next call, must equal the
    sequential order — checked against the oracle and against the same ctx with segmentation switched off; streams differ
    in level (the guess is wrong by different amounts), one goes silent mid-way, lengths are ragged"""
    lens = [F * 1152 - 700 * s for s in range(S)]
    lens[1] = 17 * 1152 + 3
    pcm = np.zeros((S, 2, F * 1152), np.int16)
    clips = []
    for s, n in enumerate(lens):
        c = clip(pkg, n, 1500 + 3 * s)
        c = np.clip(c.astype(np.int32) * (1 + s % 5) // 2, -32768, 32767).astype(np.int16)
        if s == 2:
            c[:, n // 2:] = 0
        clips.append(c)
        pcm[s, :, :n] = c
    enc = pkg.Encoder(44100, 2, 128, max_streams=S, max_frames=chunk)
    got = enc.encode_streams(pcm, chunk_frames=chunk, n_samples=lens)
    enc.set_rate_loop_segments(False)
    plain = enc.encode_streams(pcm, chunk_frames=chunk, n_samples=lens)
    assert got == plain
    for s in (0, 1, 2, S - 1):
        assert got[s] == oracle_bytes(clips[s]), s
