"""TEST INFRASTRUCTURE — drives the UNMODIFIED reference (oracle/_ref/libref.so, compiled from
/root/reference/src by oracle/Makefile) through the exact call sequence of its frame loop
(/root/reference/src/musicin.c:708-786) and dumps every hot-path intermediate.

The reference keeps all encoder state in function statics, so ONE PROCESS PER STREAM: use
`run_ref_stream()` which forks a fresh interpreter-level child for each stream.

Only tests/, tools/ fixture generators and bench.py's cpu_baseline/reference arm may import this.
"""
import ctypes as C
import multiprocessing as mp
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBREF = os.path.join(HERE, "_ref", "libref.so")


def have_ref():
    return os.path.exists(LIBREF)


class GrInfo(C.Structure):  # l3side.h:41-72
    _fields_ = [("part2_3_length", C.c_uint), ("big_values", C.c_uint), ("count1", C.c_uint),
                ("global_gain", C.c_uint), ("scalefac_compress", C.c_uint),
                ("window_switching_flag", C.c_uint), ("block_type", C.c_uint),
                ("mixed_block_flag", C.c_uint), ("table_select", C.c_uint * 3),
                ("subblock_gain", C.c_int * 3), ("region0_count", C.c_uint),
                ("region1_count", C.c_uint), ("preflag", C.c_uint), ("scalefac_scale", C.c_uint),
                ("count1table_select", C.c_uint), ("part2_length", C.c_uint), ("sfb_lmax", C.c_uint),
                ("sfb_smax", C.c_uint), ("address1", C.c_uint), ("address2", C.c_uint),
                ("address3", C.c_uint), ("quantizerStepSize", C.c_double),
                ("sfb_partition_table", C.c_void_p), ("slen", C.c_uint * 4)]


GR_FIELDS = ["part2_3_length", "big_values", "count1", "global_gain", "scalefac_compress",
             "window_switching_flag", "block_type", "mixed_block_flag", "table_select0",
             "table_select1", "table_select2", "region0_count", "region1_count", "preflag",
             "scalefac_scale", "count1table_select", "part2_length", "address1", "address2", "address3"]


def gr_to_row(g):
    return [g.part2_3_length, g.big_values, g.count1, g.global_gain, g.scalefac_compress,
            g.window_switching_flag, g.block_type, g.mixed_block_flag, g.table_select[0],
            g.table_select[1], g.table_select[2], g.region0_count, g.region1_count, g.preflag,
            g.scalefac_scale, g.count1table_select, g.part2_length, g.address1, g.address2, g.address3]


class GrCh(C.Structure):
    _fields_ = [("tt", GrInfo)]


class GrPair(C.Structure):
    _fields_ = [("ch", GrCh * 2)]


class SideInfo(C.Structure):  # l3side.h:74-85
    _fields_ = [("main_data_begin", C.c_int), ("private_bits", C.c_uint), ("resvDrain", C.c_int),
                ("scfsi", (C.c_uint * 4) * 2), ("gr", GrPair * 2)]


class PsyRatio(C.Structure):  # l3side.h:36-39
    _fields_ = [("l", ((C.c_double * 21) * 2) * 2), ("s", (((C.c_double * 3) * 12) * 2) * 2)]


class Scalefac(C.Structure):  # l3side.h:101-104
    _fields_ = [("l", ((C.c_int * 22) * 2) * 2), ("s", (((C.c_int * 3) * 13) * 2) * 2)]


class Layer(C.Structure):  # common.h:285-298
    _fields_ = [(n, C.c_int) for n in ("version", "lay", "error_protection", "bitrate_index",
                                       "sampling_frequency", "padding", "extension", "mode",
                                       "mode_ext", "copyright", "original", "emphasis")]


class FrameParams(C.Structure):  # common.h:302-310
    _fields_ = [("header", C.POINTER(Layer)), ("actual_mode", C.c_int), ("alloc", C.c_void_p),
                ("tab_num", C.c_int), ("stereo", C.c_int), ("jsbound", C.c_int), ("sblimit", C.c_int)]


BITRATES_L3 = [0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320]  # common.c:125
SFREQ_IDX = {44100: 0, 48000: 1, 32000: 2}  # common.c:115 (MPEG-1 row)


def frame_geometry(sfreq, n_ch, bitrate):
    """musicin.c:562-572,729-746: slots per frame (never padded), bits per frame, mean_bits."""
    avg = (1152.0 / (sfreq / 1000.0)) * (bitrate / 8.0)
    whole = int(avg)
    bits_per_frame = 8 * whole
    sideinfo_len = 32 + (136 if n_ch == 1 else 256)
    mean_bits = (bits_per_frame - sideinfo_len) // 2
    return whole, bits_per_frame, mean_bits


class RefEncoder:
    """One per PROCESS (reference statics)."""

    def __init__(self, sfreq=44100, n_ch=2, bitrate=128, mp3_path=None):
        self.lib = lib = C.CDLL(LIBREF)
        self.sfreq, self.n_ch, self.bitrate = sfreq, n_ch, bitrate
        self.info = Layer(version=1, lay=3, error_protection=0, bitrate_index=BITRATES_L3.index(bitrate),
                          sampling_frequency=SFREQ_IDX[sfreq], padding=0, extension=0,
                          mode=(3 if n_ch == 1 else 0), mode_ext=0, copyright=0, original=0, emphasis=0)
        self.fr_ps = FrameParams(header=C.pointer(self.info), tab_num=-1, alloc=None)
        lib.hdr_to_frps(C.byref(self.fr_ps))
        self.whole_spf, self.bits_per_frame, self.mean_bits = frame_geometry(sfreq, n_ch, bitrate)
        self.side = SideInfo()
        self.ratio = PsyRatio()
        self.scalefac = Scalefac()
        self.buffer = ((C.c_short * 1152) * 2)()
        self.sam = ((C.c_short * 1344) * 2)()
        self.win_que = ((C.c_double * 512) * 2)()
        self.sbs = ((((C.c_double * 32) * 18) * 3) * 2)()  # L3SBS [ch][3][18][32]
        self.xr = (((C.c_double * 576) * 2) * 2)()
        self.xr_dec = (((C.c_double * 576) * 2) * 2)()
        self.pe = ((C.c_double * 2) * 2)()
        self.l3_enc = (((C.c_int * 576) * 2) * 2)()
        self.snr32 = (C.c_float * 32)()
        self.bs = C.create_string_buffer(256)
        self.mp3_path = mp3_path
        self._tmp = None
        if mp3_path is None:
            self._tmp = tempfile.NamedTemporaryFile(suffix=".mp3", delete=False)
            self._tmp.close()
            self.mp3_path = self._tmp.name
        lib.open_bit_stream_w(self.bs, self.mp3_path.encode(), 4096)
        lib.L3psycho_anal.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_double,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.iteration_loop.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.III_format_bitstream.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.mdct_sub.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        lib.window_subband.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        lib.filter_subband.argtypes = [C.c_void_p, C.c_void_p]

    def encode_frame(self, pcm):
        """pcm: int16 [n_ch][1152]. Returns dict of per-frame arrays (musicin.c:751-786 order)."""
        lib, n_ch = self.lib, self.n_ch
        buf = np.ctypeslib.as_array(self.buffer)
        buf[:] = 0
        buf[:n_ch] = pcm
        out = {}
        for gr in range(2):
            for ch in range(n_ch):
                lib.L3psycho_anal(C.addressof(self.buffer[ch]) + 2 * 576 * gr, self.sam[ch], ch, 3, self.snr32,
                                  float(self.sfreq), self.ratio.l[gr][ch], self.ratio.s[gr][ch],
                                  C.addressof(self.pe[gr]) + 8 * ch, C.byref(self.side.gr[gr].ch[ch].tt))
        out["pe"] = np.ctypeslib.as_array(self.pe).copy()[:, :n_ch]
        out["ratio_l"] = np.ctypeslib.as_array(self.ratio.l).copy()[:, :n_ch]
        out["ratio_s"] = np.ctypeslib.as_array(self.ratio.s).copy()[:, :n_ch]
        out["block_type"] = np.array([[self.side.gr[g].ch[c].tt.block_type for c in range(n_ch)] for g in range(2)],
                                     dtype=np.int32)
        win_buf = (C.c_void_p * 2)(C.addressof(self.buffer[0]), C.addressof(self.buffer[1]))
        for gr in range(2):
            for ch in range(n_ch):
                for j in range(18):
                    lib.window_subband(C.addressof(win_buf) + 8 * ch, self.win_que[ch], ch)
                    lib.filter_subband(self.win_que[ch], self.sbs[ch][gr + 1][j])
        sbs = np.ctypeslib.as_array(self.sbs)  # [ch][3][18][32]
        out["sb"] = np.stack([sbs[:n_ch, 1], sbs[:n_ch, 2]], axis=0).copy()  # [gr][ch][18][32]
        lib.mdct_sub(self.sbs, self.xr, n_ch, C.byref(self.side), 2)
        out["xr"] = np.ctypeslib.as_array(self.xr).copy()[:, :n_ch]
        lib.iteration_loop(self.pe, self.xr, C.byref(self.ratio), C.byref(self.side), self.l3_enc,
                           self.mean_bits, n_ch, self.xr_dec, C.byref(self.scalefac), C.byref(self.fr_ps),
                           0, self.bits_per_frame)
        out["ix"] = np.ctypeslib.as_array(self.l3_enc).copy()[:, :n_ch]
        out["gi"] = np.array([[gr_to_row(self.side.gr[g].ch[c].tt) for c in range(n_ch)] for g in range(2)],
                             dtype=np.int32)
        out["qstep"] = np.array([[self.side.gr[g].ch[c].tt.quantizerStepSize for c in range(n_ch)]
                                 for g in range(2)])
        out["scalefac_l"] = np.ctypeslib.as_array(self.scalefac.l).copy()[:, :n_ch]
        out["scalefac_s"] = np.ctypeslib.as_array(self.scalefac.s).copy()[:, :n_ch]
        out["scfsi"] = np.ctypeslib.as_array(self.side.scfsi).copy()[:n_ch]
        out["resv_drain"] = np.int32(self.side.resvDrain)
        out["main_data_begin_in"] = np.int32(self.side.main_data_begin)
        lib.III_format_bitstream(self.bits_per_frame, C.byref(self.fr_ps), self.l3_enc, C.byref(self.side),
                                 C.byref(self.scalefac), self.bs, self.xr, None, 0)
        out["main_data_begin_next"] = np.int32(self.side.main_data_begin)
        return out

    def finish(self):
        self.lib.III_FlushBitstream()
        self.lib.close_bit_stream_w(self.bs)
        with open(self.mp3_path, "rb") as f:
            data = f.read()
        if self._tmp is not None:
            os.unlink(self.mp3_path)
        return data


def _child(conn, pcm, sfreq, n_ch, bitrate, keys):
    try:
        enc = RefEncoder(sfreq, n_ch, bitrate)
        n_frames = (pcm.shape[1] + 1151) // 1152
        padded = np.zeros((n_ch, n_frames * 1152), dtype=np.int16)
        padded[:, :pcm.shape[1]] = pcm
        frames = []
        for f in range(n_frames):
            o = enc.encode_frame(padded[:, f * 1152:(f + 1) * 1152])
            frames.append({k: v for k, v in o.items() if keys is None or k in keys})
        mp3 = enc.finish()
        res = {k: np.stack([fr[k] for fr in frames]) for k in frames[0]}
        res["mp3"] = np.frombuffer(mp3, dtype=np.uint8)
        conn.send(res)
    except Exception as e:  # pragma: no cover
        conn.send(e)
    finally:
        conn.close()


def run_ref_stream(pcm, sfreq=44100, bitrate=128, keys=None):
    """Encode one stream with the real reference in a fresh process.
    pcm: int16 [n_ch][n_samples] (planar). Returns dict of arrays stacked over frames + 'mp3' bytes."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    n_ch = pcm.shape[0]
    ctx = mp.get_context("fork")
    parent, child = ctx.Pipe()
    p = ctx.Process(target=_child, args=(child, pcm, sfreq, n_ch, bitrate, keys))
    p.start()
    child.close()
    res = None
    try:
        # the reference abort()s / exit()s on conditions it cannot handle (e.g. loop.c:358): never block on it
        while True:
            if parent.poll(0.2):
                res = parent.recv()
                break
            if not p.is_alive():
                if parent.poll(0.2):
                    res = parent.recv()
                break
    except EOFError:
        res = None
    p.join(timeout=5)
    if res is None:
        raise RuntimeError("reference process died (exit code %s) - it abort()s on inputs it cannot encode" % p.exitcode)
    if isinstance(res, Exception):
        raise res
    return res
