"""CPU: the test decoder (oracle/mp3dec.py, ISO 11172-3 §2.4.3.4 in numpy) against the reference's own outputs:
decoding the byte streams the unmodified reference CLI wrote must give back exactly the quantised spectra the
reference's iteration_loop produced (tests/golden ix), and a time signal that matches the encoder input at the
codec delay.  This pins the decoder that the GPU tests use for the decoded-SNR report."""
import os

import numpy as np
import pytest

import mp3dec
from util import ROOT

# minimum decoded SNR in dB.  The full-scale tone (scfsi case) is exempt: the reference's quantiser table ends at 2047
# (pow_nint.h:16-50), the tone's peak lines saturate there (0.35 % of all values are exactly 2047) and the reference's
# own stream decodes with < 1 dB SNR - a property of the reference that the GPU path reproduces bit for bit.
CASES = {"cfg1_44k_stereo_128": 12.0, "cfg2_32k_mono_64": 8.0, "cfg3_48k_stereo_320": 35.0, "loud_44k_stereo_128": 8.0,
         "scfsi_44k_stereo_128": -1.0}


@pytest.mark.parametrize("name", sorted(CASES))
def test_decoder_recovers_reference_spectra_and_signal(golden, name):
    g = golden[name]
    data = open(os.path.join(ROOT, "tests", "golden", "cli_%s.mp3" % name), "rb").read()
    n_ch = g["pcm"].shape[0]
    sfreq, _, xr, ix, bt, ok = mp3dec.decode_spectra(data)
    assert sfreq == int(g["sfreq"]) and ok.all()
    ref_ix = np.ascontiguousarray(g["ix"][:, :, :n_ch]).reshape(-1, n_ch, 576).astype(np.int32)
    assert np.array_equal(np.abs(ix), ref_ix)                                   # Huffman + reorder, incl. short blocks
    assert np.array_equal(bt, np.ascontiguousarray(g["gi"][:, :, :n_ch, 6]).reshape(-1, n_ch))
    pcm = mp3dec.synthesize(xr, bt)
    snr = mp3dec.snr_vs_original(g["pcm"], pcm)
    print(f"{name}: decoded SNR vs encoder input {snr:.1f} dB at delay {mp3dec.CODEC_DELAY}")
    assert snr >= CASES[name]
