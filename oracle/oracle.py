"""TEST INFRASTRUCTURE — ctypes binding of oracle/liboracle.so (the C restatement, l3oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")


class GrInfo(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("part2_3_length", "big_values", "count1", "global_gain", "scalefac_compress",
                                       "window_switching_flag", "block_type", "mixed_block_flag")] + \
               [("table_select", C.c_int * 3)] + \
               [(n, C.c_int) for n in ("region0_count", "region1_count", "preflag", "scalefac_scale",
                                       "count1table_select", "part2_length", "address1", "address2", "address3")]


class Frame(C.Structure):
    _fields_ = [("sb", (((C.c_double * 32) * 18) * 2) * 2), ("xr", ((C.c_double * 576) * 2) * 2),
                ("pe", (C.c_double * 2) * 2), ("ratio_l", ((C.c_double * 21) * 2) * 2),
                ("ratio_s", (((C.c_double * 3) * 12) * 2) * 2), ("block_type", (C.c_int * 2) * 2),
                ("max_bits", (C.c_int * 2) * 2), ("ix", ((C.c_int * 576) * 2) * 2), ("gi", (GrInfo * 2) * 2),
                ("qstep", (C.c_double * 2) * 2), ("scalefac_l", ((C.c_int * 22) * 2) * 2),
                ("scalefac_s", (((C.c_int * 3) * 13) * 2) * 2), ("scfsi", (C.c_int * 4) * 2),
                ("resv_drain", C.c_int), ("resv_size", C.c_int)]


FRAME_DT = np.dtype([("sb", "f8", (2, 2, 18, 32)), ("xr", "f8", (2, 2, 576)), ("pe", "f8", (2, 2)),
                     ("ratio_l", "f8", (2, 2, 21)), ("ratio_s", "f8", (2, 2, 12, 3)), ("block_type", "i4", (2, 2)),
                     ("max_bits", "i4", (2, 2)), ("ix", "i4", (2, 2, 576)), ("gi", "i4", (2, 2, 20)),
                     ("qstep", "f8", (2, 2)), ("scalefac_l", "i4", (2, 2, 22)), ("scalefac_s", "i4", (2, 2, 13, 3)),
                     ("scfsi", "i4", (2, 4)), ("resv_drain", "i4"), ("resv_size", "i4")], align=True)
assert FRAME_DT.itemsize == C.sizeof(Frame), (FRAME_DT.itemsize, C.sizeof(Frame))

_lib = None


def build():
    subprocess.run(["make", "-s", "-C", HERE, "liboracle.so"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.l3o_encode_stream.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_long]
        L.l3o_polyphase.argtypes = [C.c_void_p, C.c_long, C.c_void_p]
        L.l3o_mdct_granule.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.l3o_fft.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.l3o_quantize_count.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.l3o_count_bits.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.l3o_format_stream.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p]
        L.l3o_format_stream.restype = C.c_long
        _lib = L
    return _lib


def encode_stream(pcm, sfreq=44100, bitrate=128):
    """pcm int16 [n_ch][n]; returns a structured array of frames (FRAME_DT), fields indexed [frame][gr][ch]...
    Channel axis always has size 2 (mono leaves ch 1 zero), like the reference's own arrays."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    n_ch, n = pcm.shape
    n_frames = (n + 1151) // 1152
    out = np.zeros(n_frames, dtype=FRAME_DT)
    rc = lib().l3o_encode_stream(sfreq, n_ch, bitrate, pcm.ctypes.data, n, out.ctypes.data, n_frames)
    assert rc == 0
    return out


def polyphase(pcm_1ch):
    pcm = np.ascontiguousarray(pcm_1ch, dtype=np.int16)
    n_slots = pcm.shape[0] // 32
    sb = np.zeros((n_slots, 32))
    lib().l3o_polyphase(pcm.ctypes.data, n_slots, sb.ctypes.data)
    return sb


def mdct_granule(prev, cur, block_type):
    prev = np.ascontiguousarray(prev, dtype=np.float64)
    cur = np.ascontiguousarray(cur, dtype=np.float64)
    xr = np.zeros(576)
    lib().l3o_mdct_granule(prev.ctypes.data, cur.ctypes.data, int(block_type), xr.ctypes.data)
    return xr


def fft(x):
    x = np.array(x, dtype=np.float32)
    n = x.shape[0]
    e = np.zeros(n // 2 + 1, dtype=np.float32)
    p = np.zeros(n // 2 + 1, dtype=np.float32)
    lib().l3o_fft(x.ctypes.data, n, e.ctypes.data, p.ctypes.data)
    return e, p


def gi_row(g):
    return [g.part2_3_length, g.big_values, g.count1, g.global_gain, g.scalefac_compress, g.window_switching_flag,
            g.block_type, g.mixed_block_flag, g.table_select[0], g.table_select[1], g.table_select[2],
            g.region0_count, g.region1_count, g.preflag, g.scalefac_scale, g.count1table_select, g.part2_length,
            g.address1, g.address2, g.address3]


def quantize_count(xr_abs, q, block_type, sr_idx):
    xr_abs = np.ascontiguousarray(xr_abs, dtype=np.float64)
    ix = np.zeros(576, dtype=np.int32)
    g = GrInfo()
    bits = lib().l3o_quantize_count(xr_abs.ctypes.data, int(q), int(block_type), int(sr_idx), ix.ctypes.data, C.byref(g))
    return bits, ix, np.array(gi_row(g), dtype=np.int32)


def count_bits(ix, block_type, sr_idx):
    ix = np.ascontiguousarray(ix, dtype=np.int32)
    g = GrInfo()
    bits = lib().l3o_count_bits(ix.ctypes.data, int(block_type), int(sr_idx), C.byref(g))
    return bits, np.array(gi_row(g), dtype=np.int32)


def format_stream(frames, n_ch, sfreq=44100, bitrate=128):
    """frames: FRAME_DT array from encode_stream -> (bytes the reference CLI writes, main_data_begin per frame)"""
    frames = np.ascontiguousarray(frames)
    n = frames.shape[0]
    cap = n * 1440 + 16
    out = np.zeros(cap, dtype=np.uint8)
    mdb = np.zeros(n, dtype=np.int32)
    ln = lib().l3o_format_stream(sfreq, n_ch, bitrate, frames.ctypes.data, n, out.ctypes.data, cap, mdb.ctypes.data)
    assert ln >= 0
    return out[:ln].tobytes(), mdb
