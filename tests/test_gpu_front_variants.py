"""GPU: the tolerance-path variants of the fused filterbank + MDCT kernel (mp3gpu_set_front_variant, front_fast.cuh) against
the oracle / the reference's own dumps.  Tolerances are BASELINE.json's north star, written here: FP64 path (FMA, folded MDCT)
<= 1e-12, FP32 path (Lee DCT, float spectra) <= 1e-5, both relative to the largest |value| of the granule.
Whole-pipeline effect: fraction of byte-identical frames and decodability / decoded SNR against the reference CLI's file."""
import os

import numpy as np
import pytest

import oracle
from util import ROOT, oracle_flat, pad_frames, psy_array

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

CASES = ["cfg1_44k_stereo_128", "cfg2_32k_mono_64", "cfg3_48k_stereo_320", "loud_44k_stereo_128", "scfsi_44k_stereo_128"]
TOL = {"fma": 1e-12, "fma_tc": 1e-12, "fp32": 1e-5}


def rel_err(a, b):
    scale = np.maximum(np.abs(b).max(axis=-1, keepdims=True), 1e-300)
    return (np.abs(a - b) / scale).max()


@pytest.mark.parametrize("variant", ["fma", "fma_tc", "fp32"])
@pytest.mark.parametrize("name", CASES)
def test_xr_within_tolerance(pkg, golden, name, variant):
    g = golden[name]
    pcm, fs, br = g["pcm"], int(g["sfreq"]), int(g["bitrate"])
    n_ch = pcm.shape[0]
    padded, nf = pad_frames(pcm)
    o = oracle_flat(oracle.encode_stream(pcm, fs, br), n_ch)
    dev = torch.device("cuda", 0)
    enc = pkg.Encoder(fs, n_ch, br, max_streams=1, max_frames=nf)
    enc.set_front_variant(variant)
    assert enc.front_variant_info()["name"] == variant
    psy = pkg.host.psy_from_numpy(psy_array(o)[None], dev)
    xr = enc.subband_mdct_batch(torch.from_numpy(padded[None].copy()).to(dev), psy).cpu().numpy()[0]
    e = rel_err(xr, o["xr"])
    nh = len(g["xr_head"])
    ref = np.ascontiguousarray(g["xr_head"][:, :, :n_ch]).reshape(nh * 2 * n_ch, 576)
    e_ref = rel_err(xr[:len(ref)], ref)
    print(f"{name} / {variant}: max error relative to the granule maximum {e:.2e} vs oracle, {e_ref:.2e} vs the reference dump "
          f"(block types present: {sorted(set(o['block_type'].tolist()))})")
    assert e <= TOL[variant] and e_ref <= TOL[variant]


@pytest.mark.parametrize("variant", ["fma", "fma_tc", "fp32"])
@pytest.mark.parametrize("name", CASES)
def test_whole_pipeline_against_reference_cli(pkg, golden, name, variant):
    """byte stream with a tolerance-path front end vs the file the reference CLI wrote: identical-frame fraction (1.0
    expected for FMA: an error of 1e-15 flips a quantised value only on an exact rounding boundary), and the stream must
    stay a valid Layer III stream with the reference's decoded SNR to within 0.05 dB"""
    import mp3dec
    g = golden[name]
    pcm, fs, br = g["pcm"], int(g["sfreq"]), int(g["bitrate"])
    nf = (pcm.shape[1] + 1151) // 1152
    enc = pkg.Encoder(fs, pcm.shape[0], br, max_streams=1, max_frames=nf)
    enc.set_front_variant(variant)
    got = enc.encode_streams(pcm[None])[0]
    ref = open(os.path.join(ROOT, "tests", "golden", "cli_%s.mp3" % name), "rb").read()[:-1]
    FB = enc.frame_bytes
    n = (len(ref) + FB - 1) // FB
    same = sum(1 for k in range(n) if got[k * FB:(k + 1) * FB] == ref[k * FB:(k + 1) * FB])
    _, d1, ok1 = mp3dec.decode(got)
    _, d2, ok2 = mp3dec.decode(ref)
    s1, s2 = mp3dec.snr_vs_original(pcm, d1), mp3dec.snr_vs_original(pcm, d2)
    print(f"{name} / {variant}: {same}/{n} frames byte-identical to the reference CLI; decoded SNR {s1:.3f} dB (reference {s2:.3f} dB)")
    assert ok1.all() and abs(len(got) - len(ref)) <= FB
    assert abs(s1 - s2) <= 0.05
    if variant != "fp32":
        assert same == n


def test_variant_is_validated(pkg):
    enc = pkg.Encoder(44100, 2, 128)
    assert enc.lib.mp3gpu_set_front_variant(enc.ctx, 7) == -1
    assert enc.front_variant_info() == {"name": "exact", "bytes_per_gc": 5764}
    enc.set_front_variant("fp32")
    assert enc.front_variant_info() == {"name": "fp32", "bytes_per_gc": 3460}
