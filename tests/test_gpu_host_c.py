"""GPU: the batched C host (examples/mp3gpu_encode.c — the reference's frame loop on libmp3gpu.so, plain C99 over the C ABI)
against the files the UNMODIFIED reference CLI wrote for the golden inputs.  Byte-identical, including the one zero byte
close_bit_stream_w() appends (common.c:968-974)."""
import os
import subprocess
import tempfile

import pytest

from util import ROOT, cli_flags, write_wav

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "examples", "mp3gpu_encode")


def build_exe():
    subprocess.run(["gcc", "-O2", "-std=c99", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "mp3gpu_encode.c"),
                    "-o", EXE, "-L" + os.path.join(ROOT, "mp3-enc-bsd_b200"), "-lmp3gpu", "-Wl,-rpath,$ORIGIN/../mp3-enc-bsd_b200"], check=True)


def run(flags, names, golden, chunk):
    with tempfile.TemporaryDirectory() as d:
        wavs = []
        for n in names:
            p = os.path.join(d, n + ".wav")
            write_wav(p, golden[n]["pcm"], int(golden[n]["sfreq"]))
            wavs.append(p)
        out = subprocess.run([EXE] + flags + ["-c", str(chunk), d] + wavs, capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        print(out.stdout.strip().splitlines()[-1])
        return [open(os.path.join(d, n + ".mp3"), "rb").read() for n in names]


@pytest.mark.parametrize("chunk", [30, 7])
def test_three_streams_of_different_length_in_one_batch(golden, chunk):
    """46-, 31- and 23-frame streams (44.1 kHz stereo 128 kbps) as ONE batch through the C host"""
    build_exe()
    names = ["cfg1_44k_stereo_128", "loud_44k_stereo_128", "scfsi_44k_stereo_128"]
    got = run(cli_flags(2, 44100, 128), names, golden, chunk)
    for n, g in zip(names, got):
        ref = open(os.path.join(ROOT, "tests", "golden", "cli_%s.mp3" % n), "rb").read()
        assert g == ref, (n, len(g), len(ref))


@pytest.mark.parametrize("name", ["cfg2_32k_mono_64", "cfg3_48k_stereo_320"])
def test_other_configurations(golden, name):
    build_exe()
    g = golden[name]
    got = run(cli_flags(g["pcm"].shape[0], int(g["sfreq"]), int(g["bitrate"])), [name], golden, 30)[0]
    ref = open(os.path.join(ROOT, "tests", "golden", "cli_%s.mp3" % name), "rb").read()
    assert got == ref, (len(got), len(ref))
