"""GPU: one long stream cut into independently encoded segments (BASELINE configs[4]) vs the whole-stream encode."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_segments_identical_when_reservoir_is_disabled(pkg):
    """48 kHz stereo 320 kbps: frame = 7680 bits => ResvMax = 0 (reservoir.c:45-91), so nothing but the signal
    history couples frames and the default 8-frame pre-roll must reproduce the whole-stream encode BYTE FOR BYTE:
    proves filterbank / MDCT / psy state (block-type state machine, pre-echo history) and calc_scfsi's statics
    re-converge inside the pre-roll.  With the minimal 2-frame pre-roll only the stale-en[] quirk may differ."""
    seg = pkg.segment
    fs, br = 48000, 320
    pcm = pkg.synth.config3(2.0, fs, seeds=(4, 5))
    enc = pkg.Encoder(fs, 2, br, max_streams=1, max_frames=32)
    whole = enc.encode_streams(pcm[None])[0]
    cut = seg.encode_long_stream(pcm, 7, seg.gpu_batch_encoder(pkg, fs, 2, br, chunk_frames=8), enc.frame_bytes)
    frac, diff = seg.frame_identity(whole, cut, enc.frame_bytes)
    assert whole == cut, (len(whole), len(cut), frac, diff[:10])
    cut2 = seg.encode_long_stream(pcm, 7, seg.gpu_batch_encoder(pkg, fs, 2, br, chunk_frames=8), enc.frame_bytes, preroll_frames=2)
    frac2, diff2 = seg.frame_identity(whole, cut2, enc.frame_bytes)
    print(f"320 kbps, 7 segments: pre-roll 8 frames 100% identical, pre-roll 2 frames {100 * frac2:.1f}% (frames {diff2})")
    assert frac2 >= 0.95 and all(d % 12 == 0 for d in diff2)


def test_transient_stream_segments_with_reservoir(pkg):
    """32 kHz mono 64 kbps transient-heavy stream (short-block switching) in 5 segments: frames before the first seam
    are identical; after a seam frames differ only while the reservoir differs; the stream keeps its frame grid."""
    seg = pkg.segment
    fs, br = 32000, 64
    pcm = pkg.synth.config2(6.0, fs, seed=3)
    enc = pkg.Encoder(fs, 1, br, max_streams=1, max_frames=32)
    FB = enc.frame_bytes
    whole = enc.encode_streams(pcm[None])[0]
    plan = seg.plan_segments((pcm.shape[1] + 1151) // 1152, 5, 2)
    cut = seg.encode_long_stream(pcm, 5, seg.gpu_batch_encoder(pkg, fs, 1, br, chunk_frames=16), FB, preroll_frames=2)
    frac, diff = seg.frame_identity(whole, cut, FB)
    n = (len(whole) + FB - 1) // FB
    print(f"{n} frames, 5 segments: {100 * frac:.1f}% byte-identical to the whole-stream encode; first differing frame {diff[:1]}")
    assert abs(len(cut) - len(whole)) < FB
    assert all(cut[k * FB:k * FB + 2] == whole[:2] for k in range(n))              # sync words on the frame grid
    # everything the first segment delivers, except its last frames whose slack the whole-stream encode fills with
    # the next frame's main data, is identical
    assert all(d >= plan[1].first_frame - 2 for d in diff), diff[:5]
    # main_data_begin of the first frame of each segment is 0
    for s in plan:
        mdb = (cut[s.first_frame * FB + 4] << 1) | (cut[s.first_frame * FB + 5] >> 7)
        assert mdb == 0


def test_many_segments_in_one_batch_match_single_segment_runs(pkg):
    """batching segments as streams changes nothing: 6 segments in one call == each segment encoded alone"""
    seg = pkg.segment
    fs, br = 44100, 128
    pcm = pkg.synth.config1(3.2, fs, seeds=(21, 22))
    FB = pkg.Encoder(fs, 2, br, max_streams=1, max_frames=1).frame_bytes
    plan = seg.plan_segments((pcm.shape[1] + 1151) // 1152, 6, 2)
    batched = seg.encode_segments(pcm, plan, seg.gpu_batch_encoder(pkg, fs, 2, br, chunk_frames=5))
    for s in plan:
        alone = seg.encode_segments(pcm, [s], seg.gpu_batch_encoder(pkg, fs, 2, br, chunk_frames=32))
        assert alone[s.index] == batched[s.index], s


def test_segmented_stream_decodes_and_matches_whole_stream_quality(pkg):
    """the stitched stream of a segmented encode is valid Layer III (every frame's main data is reachable, Huffman
    data parses, part2_3_length is consistent) and sounds like the whole-stream encode: decoded SNR against the
    encoder input within 0.5 dB, decoded-vs-decoded SNR reported"""
    import mp3dec
    seg = pkg.segment
    fs, br = 44100, 128
    pcm = pkg.synth.config1(4.0, fs, seeds=(31, 32))
    enc = pkg.Encoder(fs, 2, br, max_streams=1, max_frames=32)
    FB = enc.frame_bytes
    whole = enc.encode_streams(pcm[None])[0]
    cut = seg.encode_long_stream(pcm, 6, seg.gpu_batch_encoder(pkg, fs, 2, br, chunk_frames=16), FB)
    frac, diff = seg.frame_identity(whole, cut, FB)
    _, dw, okw = mp3dec.decode(whole)
    _, dc, okc = mp3dec.decode(cut)
    assert okw.all() and okc.all()
    sw, sc = mp3dec.snr_vs_original(pcm, dw), mp3dec.snr_vs_original(pcm, dc)
    cross = mp3dec.snr_db(dw, dc)
    print(f"6 segments: {100 * frac:.1f}% frames byte-identical; decoded SNR vs input: whole {sw:.2f} dB, segmented {sc:.2f} dB; "
          f"segmented vs whole decode {cross:.1f} dB")
    assert abs(sw - sc) < 0.5 and sw > 12.0
