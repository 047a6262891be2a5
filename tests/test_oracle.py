"""CPU: pin the oracle (oracle/l3oracle.c) against the reference's own outputs.

 * test_oracle_matches_golden: against the fixtures under tests/golden/ that tools/make_golden.py
   generated from the UNMODIFIED reference (oracle/_ref/libref.so) — runs everywhere.
 * test_oracle_matches_reference_live: against libref.so itself on fresh inputs — only where the
   reference has been compiled (the build container).
The reference ships no golden vectors / KATs of its own.
"""
import numpy as np
import pytest

import oracle
import ref_harness
import os

from util import ROOT, oracle_flat

XR_TOL = 5e-14   # the oracle uses the plain 36-term dot product for block type 0; the reference's hand-unrolled
                 # form (mdct.c:199-509) only differs in summation order (measured <= 2e-14 at |xr| <= 1.5)


def check_against_ref(o, r, n_ch):
    for k in ("pe", "ratio_l", "ratio_s"):
        assert np.array_equal(o[k][:, :, :n_ch], r[k]), k            # FP32/FP64 psy: bit-identical
    for k in ("block_type", "gi", "scalefac_l", "scalefac_s"):
        assert np.array_equal(o[k][:, :, :n_ch], r[k]), k
    assert np.array_equal(o["ix"][:, :, :n_ch], r["ix"].astype(np.int32)), "ix"
    assert np.array_equal(o["scfsi"][:, :n_ch], r["scfsi"]), "scfsi"
    assert np.array_equal(o["resv_drain"], r["resv_drain"])


@pytest.mark.parametrize("name", ["cfg1_44k_stereo_128", "cfg2_32k_mono_64", "cfg3_48k_stereo_320", "loud_44k_stereo_128",
                                  "scfsi_44k_stereo_128"])
def test_oracle_matches_golden(golden, name):
    g = golden[name]
    pcm = g["pcm"]
    n_ch = pcm.shape[0]
    o = oracle.encode_stream(pcm, int(g["sfreq"]), int(g["bitrate"]))
    check_against_ref(o, g, n_ch)
    nh = len(g["sb_head"])
    assert np.array_equal(o["sb"][:nh, :, :n_ch], g["sb_head"])        # subband samples: bit-identical
    assert np.abs(o["xr"][:nh, :, :n_ch] - g["xr_head"]).max() <= XR_TOL
    assert np.abs(o["xr"][:, :, :n_ch].sum(axis=-1) - g["xr_sum"]).max() <= 1e-11
    assert np.array_equal(o["resv_size"], g["main_data_begin_next"] * 8)  # reservoir == formatter's back pointer
    if name.startswith("scfsi"):
        assert g["scfsi"].sum() > 0


@pytest.mark.skipif(not ref_harness.have_ref(), reason="oracle/_ref/libref.so not built (reference sources absent)")
@pytest.mark.parametrize("case", ["cfg1", "cfg2", "cfg3", "square", "silence", "dc", "ragged"])
def test_oracle_matches_reference_live(pkg, case):
    s = pkg.synth
    fs, br = 44100, 128
    if case == "cfg1":
        pcm = s.config1(2.0, seeds=(11, 12))
    elif case == "cfg2":
        pcm, fs, br = s.config2(2.5, seed=13), 32000, 64
    elif case == "cfg3":
        pcm, fs, br = s.config3(1.5, seeds=(14, 15)), 48000, 320
    elif case == "square":
        t = np.arange(fs) / fs
        pcm = (np.stack([np.sign(np.sin(2 * np.pi * 200 * t))] * 2) * 32000).astype(np.int16)
    elif case == "silence":
        pcm = np.zeros((2, 5000), np.int16)
    elif case == "dc":
        pcm = np.full((2, 20000), 16000, np.int16)
    else:  # ragged: not a whole number of frames, mono, other bitrate
        pcm, fs, br = s.config1(0.777, seeds=(5, 6))[:1], 48000, 96
    n_ch = pcm.shape[0]
    r = ref_harness.run_ref_stream(pcm, fs, br)
    o = oracle.encode_stream(pcm, fs, br)
    check_against_ref(o, r, n_ch)
    assert np.array_equal(o["sb"][:, :, :n_ch], r["sb"])
    assert np.abs(o["xr"][:, :, :n_ch] - r["xr"]).max() <= XR_TOL


def test_oracle_fft_vs_numpy():
    rng = np.random.default_rng(3)
    for n in (1024, 256):
        x = (rng.standard_normal(n) * 3000).astype(np.float32)
        e, p = oracle.fft(x)
        F = np.fft.rfft(x.astype(np.float64))
        ref = np.abs(F) ** 2
        assert np.abs(e - ref).max() <= 2e-6 * ref.max()
        big = ref > 1e-3 * ref.max()
        d = np.angle(np.exp(1j * (p - np.angle(F))))
        assert np.abs(d[big]).max() < 1e-3


def test_stage_functions_consistent():
    """stateless stage entry points agree with the streaming encoder"""
    import mp3gpu_pkg
    s = mp3gpu_pkg.load().synth
    pcm = s.config1(0.3)
    o = oracle.encode_stream(pcm, 44100, 128)
    padded = np.zeros(len(o) * 1152, np.int16)
    padded[:pcm.shape[1]] = pcm[0]
    sb = oracle.polyphase(padded).reshape(len(o), 2, 18, 32)
    assert np.array_equal(sb, o["sb"][:, :, 0])
    xr1 = oracle.mdct_granule(sb[0, 0], sb[0, 1], o["block_type"][0, 1, 0])
    assert np.array_equal(xr1, o["xr"][0, 1, 0])
    # quantize + count at the final step reproduces the final ix / side info of a long block
    for f in range(2, len(o)):
        if o["block_type"][f, 0, 0] == 0 and o["scalefac_l"][f, 0, 0].max() == 0 and o["gi"][f, 0, 0][13] == 0:
            bits, ix, gi = oracle.quantize_count(np.abs(o["xr"][f, 0, 0]), o["gi"][f, 0, 0][3] - 210, 0, 1)
            assert np.array_equal(ix, o["ix"][f, 0, 0])
            for col in (1, 2, 8, 9, 10, 11, 12, 15):
                assert gi[col] == o["gi"][f, 0, 0][col]
            break


@pytest.mark.parametrize("name", ["cfg1_44k_stereo_128", "cfg2_32k_mono_64", "cfg3_48k_stereo_320", "loud_44k_stereo_128",
                                  "scfsi_44k_stereo_128"])
def test_oracle_formatter_matches_reference_cli_bytes(golden, name):
    """the oracle's sequential bitstream formatter (l3bitstream.c + formatBitstream.c restated) reproduces the byte
    stream the unmodified reference CLI wrote (tests/golden/cli_*.mp3 == the npz's `mp3`), incl. the back pointers"""
    g = golden[name]
    pcm, fs, br = g["pcm"], int(g["sfreq"]), int(g["bitrate"])
    fr = oracle.encode_stream(pcm, fs, br)
    data, mdb = oracle.format_stream(fr, pcm.shape[0], fs, br)
    ref = open(os.path.join(ROOT, "tests", "golden", "cli_%s.mp3" % name), "rb").read()
    assert data == ref and data == bytes(g["mp3"])
    assert np.array_equal(mdb[1:], g["main_data_begin_next"][:-1]) and mdb[0] == 0
