"""GPU: the device bitstream formatter (mp3gpu_encode_frames_mp3 / mp3gpu_format_bitstream_batch / mp3gpu_flush_mp3)
against (a) the byte streams the UNMODIFIED reference CLI wrote for the golden inputs (tests/golden/cli_*.mp3),
(b) the oracle's sequential formatter on seeded inputs, (c) the reference CLI itself when oracle/_ref/encode is present.
Byte-exact in every case.  The reference file has ONE more byte than the stream: close_bit_stream_w() (common.c:968-974)
writes the partially filled buffer byte, always zero."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oracle
from util import ROOT, cli_flags, pad_frames, write_wav

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

GOLDEN = ["cfg1_44k_stereo_128", "cfg2_32k_mono_64", "cfg3_48k_stereo_320", "loud_44k_stereo_128", "scfsi_44k_stereo_128"]


def first_diff(a, b):
    n = min(len(a), len(b))
    return next((i for i in range(n) if a[i] != b[i]), None if len(a) == len(b) else n)


@pytest.mark.parametrize("chunk", [None, 7, 1])
@pytest.mark.parametrize("name", GOLDEN)
def test_mp3_bytes_vs_reference_cli_golden(pkg, golden, name, chunk):
    g = golden[name]
    pcm, fs, br = g["pcm"], int(g["sfreq"]), int(g["bitrate"])
    nf = (pcm.shape[1] + 1151) // 1152
    enc = pkg.Encoder(fs, pcm.shape[0], br, max_streams=1, max_frames=chunk or nf)
    got = enc.encode_streams(pcm[None], chunk_frames=chunk)[0]
    ref = open(os.path.join(ROOT, "tests", "golden", "cli_%s.mp3" % name), "rb").read()
    assert ref[-1] == 0
    assert len(got) == len(ref) - 1 and got == ref[:-1], (len(got), len(ref), first_diff(got, ref))
    frames = len(ref) // enc.frame_bytes
    print(f"{name}: {frames}/{frames} frames byte-identical to the reference CLI ({len(got)} bytes, chunk={chunk})")


@pytest.mark.parametrize("pipelined", [False, True])
def test_batch_of_streams_vs_oracle_formatter(pkg, pipelined):
    """8 different streams (one silent, one going silent, one loud) in one batch, ragged chunks; in-order and pipelined
    delivery of the host bytes (mp3gpu_set_host_delivery) must give the same file"""
    S, F = 8, 20
    sy = pkg.synth
    pcm = np.stack([sy.config1(F * 1152 / 44100.0 + 0.01, seeds=(300 + 2 * i, 301 + 2 * i))[:, :F * 1152] for i in range(S)])
    pcm[2] = 0
    pcm[4, :, 9000:] = 0
    pcm[6] = np.clip(pcm[6].astype(np.int32) * 4, -32768, 32767).astype(np.int16)
    enc = pkg.Encoder(44100, 2, 128, max_streams=S, max_frames=6)
    enc.set_host_delivery(pipelined)
    mp3 = np.zeros((S, F * enc.frame_bytes + 64), np.uint8)
    f0 = 0
    for c in [6, 1, 4, 2, 6, 1]:
        enc.encode_frames_mp3(np.ascontiguousarray(pcm[:, :, f0 * 1152:(f0 + c) * 1152]), mp3)
        f0 += c
    assert f0 == F
    lengths = enc.flush_mp3(mp3, S)
    for s in range(S):
        ref, _ = oracle.format_stream(oracle.encode_stream(pcm[s], 44100, 128), 2, 44100, 128)
        got = mp3[s, :lengths[s]].tobytes()
        assert got == ref[:-1], (s, len(got), len(ref), first_diff(got, ref))
        assert not mp3[s, lengths[s]:].any()


@pytest.mark.parametrize("fs,n_ch,br", [(48000, 1, 32), (32000, 2, 320), (44100, 1, 128), (32000, 2, 64)])
def test_other_formats_vs_oracle_formatter(pkg, fs, n_ch, br):
    """small frames (96 bytes: the back pointer spans 7 frames), huge frames (reservoir disabled), transient-heavy input"""
    F = 30
    pcm = pkg.synth.config2(F * 1152 / fs + 0.01, fs, seed=11)[:, :F * 1152] if n_ch == 1 else \
        pkg.synth.config3(F * 1152 / fs + 0.01, fs, seeds=(12, 13))[:, :F * 1152]
    enc = pkg.Encoder(fs, n_ch, br, max_streams=1, max_frames=4)
    got = enc.encode_streams(pcm[None], chunk_frames=4)[0]
    ref, mdb = oracle.format_stream(oracle.encode_stream(pcm, fs, br), n_ch, fs, br)
    assert got == ref[:-1], (len(got), len(ref), first_diff(got, ref))
    print(f"{fs} Hz {n_ch} ch {br} kbps: frame {enc.frame_bytes} B, max main_data_begin {mdb.max()} B")


def test_format_bitstream_batch_stage_entry(pkg):
    """III_format_bitstream batched, device tensors, fed with what mp3gpu_encode_frames_dev produced"""
    S, F = 3, 9
    pcm = np.stack([pkg.synth.config1(F * 1152 / 44100.0 + 0.01, seeds=(50 + s, 60 + s))[:, :F * 1152] for s in range(S)])
    enc = pkg.Encoder(44100, 2, 128, max_streams=S, max_frames=F)
    dev = torch.device("cuda", 0)
    out = enc.encode_frames_dev(torch.from_numpy(pcm).to(dev))
    mp3 = torch.zeros((S, F * enc.frame_bytes), dtype=torch.uint8, device=dev)
    enc.format_bitstream_batch(out, mp3)
    lengths = enc.flush_mp3(mp3, S)
    host = mp3.cpu().numpy()
    for s in range(S):
        ref, _ = oracle.format_stream(oracle.encode_stream(pcm[s], 44100, 128), 2, 44100, 128)
        assert host[s, :lengths[s]].tobytes() == ref[:-1]


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "encode")), reason="reference CLI not built")
def test_mp3_bytes_vs_reference_cli_live(pkg):
    """a seeded clip the goldens do not contain, through the unmodified reference CLI on this box"""
    pcm = pkg.synth.config1(3.0, 44100, seeds=(77, 78))
    with tempfile.TemporaryDirectory() as tmp:
        wav, out = os.path.join(tmp, "in.wav"), os.path.join(tmp, "out.mp3")
        write_wav(wav, pcm, 44100)
        subprocess.run([os.path.join(ROOT, "oracle", "_ref", "encode")] + cli_flags(2, 44100, 128) + [wav, out], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        ref = open(out, "rb").read()
    enc = pkg.Encoder(44100, 2, 128, max_streams=1, max_frames=16)
    got = enc.encode_streams(pcm[None])[0]
    assert got == ref[:-1], (len(got), len(ref), first_diff(got, ref))


def test_interleaved_ingest_and_file_batch(pkg, tmp_path):
    """WAV files in, MP3 byte streams out (interleaved PCM de-interleaved on the device, get_audio() encode.c:256-269):
    three files of different length and a raw (header-less) file in one call; each equals the oracle's formatter output
    and, where the reference CLI is present, the CLI's file."""
    host = pkg.host
    clips = [pkg.synth.config1(1.0 + 0.37 * i, 44100, seeds=(90 + i, 95 + i)) for i in range(3)]
    clips[1] = clips[1][:, :clips[0].shape[1]]                    # two files with the same frame count share a batch
    paths = []
    for i, c in enumerate(clips):
        p = str(tmp_path / f"clip{i}.wav")
        write_wav(p, c, 44100)
        paths.append(p)
    raw = str(tmp_path / "clip_raw.pcm")
    np.ascontiguousarray(clips[2].T).astype("<i2").tofile(raw)     # raw is little-endian in effect, see host.read_pcm_file
    paths.append(raw)
    got = host.encode_files(paths, 44100, 2, 128, chunk_frames=8)
    for i, c in enumerate(clips + [clips[2]]):
        ref, _ = oracle.format_stream(oracle.encode_stream(c, 44100, 128), 2, 44100, 128)
        assert got[i] == ref[:-1], (i, len(got[i]), len(ref), first_diff(got[i], ref))
    cli = os.path.join(ROOT, "oracle", "_ref", "encode")
    if os.path.exists(cli):
        out = str(tmp_path / "ref.mp3")
        subprocess.run([cli] + cli_flags(2, 44100, 128) + [paths[3], out], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        assert open(out, "rb").read()[:-1] == got[3]
