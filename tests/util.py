"""Shared helpers for the parity tests."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PSY_DT = np.dtype([("pe", "f8"), ("ratio_l", "f8", 21), ("ratio_s", "f8", 36), ("block_type", "i4"), ("pad", "i4")])
FO_DT = np.dtype([("resv_drain", "i4"), ("main_data_begin", "i4"), ("scfsi", "u1", (2, 4))])


def oracle_flat(o, n_ch):
    """oracle frames (FRAME_DT, [frame][gr][ch]...) -> dict of arrays indexed by gc = (frame*2+gr)*n_ch+ch"""
    nf = len(o)
    ngc = nf * 2 * n_ch
    f = lambda k: np.ascontiguousarray(o[k][:, :, :n_ch]).reshape((ngc,) + o[k].shape[3:])
    d = {k: f(k) for k in ("sb", "xr", "pe", "ratio_l", "ratio_s", "block_type", "max_bits", "ix", "gi", "scalefac_l", "scalefac_s")}
    d["ratio_s"] = d["ratio_s"].reshape(ngc, 36)
    d["scfsi"] = o["scfsi"].astype(np.uint8)
    d["resv_drain"] = o["resv_drain"]
    d["resv_size"] = o["resv_size"]
    return d


def psy_array(flat):
    psy = np.zeros(len(flat["pe"]), PSY_DT)
    psy["pe"] = flat["pe"]
    psy["ratio_l"] = flat["ratio_l"]
    psy["ratio_s"] = flat["ratio_s"]
    psy["block_type"] = flat["block_type"]
    return psy


def expected_sf(flat):
    """scalefactor bytes as the kernel writes them: long l[0..21]; short s[sfb][w] at 3*sfb+w"""
    ngc = len(flat["pe"])
    sf = np.zeros((ngc, 40), np.uint8)
    for g in range(ngc):
        if flat["block_type"][g] == 2:
            sf[g, :36] = flat["scalefac_s"][g, :12].reshape(36)
        else:
            sf[g, :22] = flat["scalefac_l"][g]
    return sf


def sf_mask(block_type):
    m = np.zeros((len(block_type), 40), bool)
    m[block_type == 2, :36] = True
    m[block_type != 2, :22] = True
    return m


def pad_frames(pcm):
    n_ch, n = pcm.shape
    nf = (n + 1151) // 1152
    out = np.zeros((n_ch, nf * 1152), np.int16)
    out[:, :n] = pcm
    return out, nf


class Emul:
    """tests/emul/libemul.so: the kernel cores compiled for the host (csrc/simt.h)"""

    def __init__(self):
        self.lib = C.CDLL(os.path.join(ROOT, "tests", "emul", "libemul.so"))
        assert self.lib.emul_sizeof_psyout() == PSY_DT.itemsize
        assert self.lib.emul_sizeof_frameout() == FO_DT.itemsize

    def encode_stream(self, pcm, sfreq, bitrate):
        n_ch = pcm.shape[0]
        padded, nf = pad_frames(pcm)
        ngc = nf * 2 * n_ch
        H = 1056
        buf = np.zeros((n_ch, H + nf * 1152), np.int16)
        buf[:, H:] = padded
        r = dict(sb=np.zeros((ngc, 18, 32)), xr=np.zeros((ngc, 576)), psy=np.zeros(ngc, PSY_DT), ix=np.zeros((ngc, 576), np.int16),
                 gi=np.zeros((ngc, 20), np.int32), sf=np.zeros((ngc, 40), np.uint8), fo=np.zeros(nf, FO_DT),
                 max_bits=np.zeros(ngc, np.int32))
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = self.lib.emul_encode_stream(sfreq, n_ch, bitrate, nf, vp(buf), C.c_long(buf.shape[1]), H, vp(r["sb"]), vp(r["xr"]),
                                         vp(r["psy"]), vp(r["ix"]), vp(r["gi"]), vp(r["sf"]), vp(r["fo"]), vp(r["max_bits"]))
        assert rc == 0
        return r

    def fft(self, x):
        y = np.array(x, dtype=np.float32)
        nops, nlev = C.c_int(), C.c_int()
        rc = self.lib.emul_fft(y.ctypes.data_as(C.c_void_p), len(y), C.byref(nops), C.byref(nlev))
        return rc, y, nops.value, nlev.value


def write_wav(path, pcm, sfreq):
    """44-byte-header WAV as the reference sniffs it (musicin.c:352-368): data assumed at 0x2c, little endian"""
    import struct
    n_ch = pcm.shape[0]
    inter = np.ascontiguousarray(pcm.T).reshape(-1).astype("<i2")
    hdr = b"RIFF" + struct.pack("<I", 36 + inter.nbytes) + b"WAVEfmt " + \
        struct.pack("<IHHIIHH", 16, 1, n_ch, sfreq, sfreq * 2 * n_ch, 2 * n_ch, 16) + b"data" + struct.pack("<I", inter.nbytes)
    with open(path, "wb") as f:
        f.write(hdr + inter.tobytes())


def cli_flags(n_ch, sfreq, bitrate):
    """reference CLI flags (musicin.c:206-296) for a configuration"""
    return (["-m", "m"] if n_ch == 1 else ["-m", "s"]) + ["-s", {32000: "32", 44100: "44.1", 48000: "48"}[sfreq], "-b", str(bitrate)]


def mp3_frames(data, frame_bytes):
    return [data[i:i + frame_bytes] for i in range(0, len(data), frame_bytes)]
