"""Python host side above the C ABI (include/mp3gpu.h): a thin ctypes mirror of libmp3gpu.so.

The reference has no plugin/operator API — its boundary is the five C symbols its frame loop calls
(musicin.c:754-779).  This module gives each batched replacement a method with the same name and
argument meaning, so the parity tests read like a batched version of the reference's frame loop:

    enc = Encoder(sfreq=44100, n_ch=2, bitrate=128, max_streams=S, max_frames=F)
    psy = enc.L3psycho_anal_batch(pcm)             # l3psy.c:53
    sb  = enc.filter_subband_batch(pcm)            # encode.c:287,361
    xr  = enc.mdct_sub_batch(sb, psy)              # mdct.c:25
    out = enc.iteration_loop_batch(xr, psy)        # loop.c:232
    out = enc.encode_frames(pcm)                   # all four, fused pipeline, host buffers

torch is used only for device memory and streams (plumbing).  There is NO CPU fallback: if the
shared library is missing or no GPU is present, constructing an Encoder raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# MP3GPU_LIB selects another build of the same library (A/B runs of kernel variants); default: the in-tree build
LIB_PATH = os.environ.get("MP3GPU_LIB") or os.path.join(HERE, "libmp3gpu.so")

PSY_DT = np.dtype([("pe", "f8"), ("ratio_l", "f8", 21), ("ratio_s", "f8", 36), ("block_type", "i4"), ("pad", "i4")])
FO_DT = np.dtype([("resv_drain", "i4"), ("main_data_begin", "i4"), ("scfsi", "u1", (2, 4))])
GI_FIELDS = ["part2_3_length", "big_values", "count1", "global_gain", "scalefac_compress", "window_switching_flag",
             "block_type", "mixed_block_flag", "table_select0", "table_select1", "table_select2", "region0_count",
             "region1_count", "preflag", "scalefac_scale", "count1table_select", "part2_length", "address1",
             "address2", "address3"]

EXPORTS = ["mp3gpu_last_error", "mp3gpu_version", "mp3gpu_create", "mp3gpu_destroy", "mp3gpu_reset",
           "mp3gpu_frame_geometry", "mp3gpu_encode_frames", "mp3gpu_encode_frames_dev", "mp3gpu_sync",
           "mp3gpu_filter_subband_batch", "mp3gpu_mdct_sub_batch", "mp3gpu_subband_mdct_batch",
           "mp3gpu_L3psycho_anal_batch", "mp3gpu_iteration_loop_batch", "mp3gpu_quantize_count_batch",
           "mp3gpu_kernel_launches", "mp3gpu_profile_enable", "mp3gpu_profile_collect",
           "mp3gpu_encode_frames_mp3", "mp3gpu_encode_frames_mp3_dev", "mp3gpu_flush_mp3", "mp3gpu_flush_mp3_dev",
           "mp3gpu_frame_bytes", "mp3gpu_format_bitstream_batch", "mp3gpu_begin_segment", "mp3gpu_stream_wave", "mp3gpu_set_pcm_layout", "mp3gpu_count_bits_batch", "mp3gpu_set_host_delivery",
           "mp3gpu_reset_async", "mp3gpu_set_stream_frames", "mp3gpu_reset_streams",
           "mp3gpu_set_front_variant", "mp3gpu_get_front_variant", "mp3gpu_set_pipeline", "mp3gpu_set_psy_variant",
           "mp3gpu_set_rate_loop_segments", "mp3gpu_rate_loop_segment_stats"]
FRONT_VARIANTS = {"exact": 0, "fma": 1, "fp32": 2, "fma_tc": 3}
LEGACY_EXPORTS = ["window_subband", "filter_subband", "mdct_sub", "L3psycho_anal", "iteration_loop", "quantize", "count_bits",
                  "inner_loop", "bin_search_StepSize", "calc_runlen", "count1_bitcount", "subdivide", "bigv_tab_select",
                  "new_choose_table", "bigv_bitcount",
                  "mp3gpu_legacy_reset", "mp3gpu_legacy_kernel_launches"]
KERNEL_NAMES = ["psy_front", "psy_scan", "front_polyphase_mdct", "rate_loop", "bitstream"]


class Config(C.Structure):
    _fields_ = [("sfreq_hz", C.c_int), ("n_ch", C.c_int), ("bitrate_kbps", C.c_int), ("max_streams", C.c_int),
                ("max_frames", C.c_int), ("device", C.c_int)]


class Mp3GpuError(RuntimeError):
    pass


_lib = None


def load_library():
    """dlopen libmp3gpu.so (built in-tree by __graft_entry__.build()). Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Mp3GpuError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback for the hot path)")
        lib = C.CDLL(LIB_PATH)
        lib.mp3gpu_last_error.restype = C.c_char_p
        lib.mp3gpu_version.restype = C.c_char_p
        lib.mp3gpu_kernel_launches.restype = C.c_long
        lib.mp3gpu_kernel_launches.argtypes = [C.c_void_p]
        lib.mp3gpu_profile_enable.argtypes = [C.c_void_p, C.c_int]
        lib.mp3gpu_profile_collect.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_long), C.c_int]
        lib.mp3gpu_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        lib.mp3gpu_destroy.argtypes = [C.c_void_p]
        lib.mp3gpu_reset.argtypes = [C.c_void_p]
        lib.mp3gpu_reset_async.argtypes = [C.c_void_p, C.c_void_p]
        lib.mp3gpu_set_stream_frames.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_long), C.c_void_p]
        lib.mp3gpu_reset_streams.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        lib.mp3gpu_set_front_variant.argtypes = [C.c_void_p, C.c_int]
        lib.mp3gpu_set_psy_variant.argtypes = [C.c_void_p, C.c_int]
        lib.mp3gpu_set_pipeline.argtypes = [C.c_void_p, C.c_int]
        lib.mp3gpu_set_rate_loop_segments.argtypes = [C.c_void_p, C.c_int]
        lib.mp3gpu_rate_loop_segment_stats.argtypes = [C.c_void_p, C.POINTER(C.c_long), C.c_int]
        lib.mp3gpu_get_front_variant.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.mp3gpu_sync.argtypes = [C.c_void_p, C.c_void_p]
        lib.mp3gpu_frame_geometry.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        vp = C.c_void_p
        for name in ("mp3gpu_encode_frames", "mp3gpu_encode_frames_dev"):
            getattr(lib, name).argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]
        lib.mp3gpu_filter_subband_batch.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp]
        lib.mp3gpu_mdct_sub_batch.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, vp]
        lib.mp3gpu_subband_mdct_batch.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, vp]
        lib.mp3gpu_L3psycho_anal_batch.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp]
        lib.mp3gpu_iteration_loop_batch.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]
        lib.mp3gpu_quantize_count_batch.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp, vp, vp]
        for name in ("mp3gpu_encode_frames_mp3", "mp3gpu_encode_frames_mp3_dev"):
            getattr(lib, name).argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_long, vp]
        for name in ("mp3gpu_flush_mp3", "mp3gpu_flush_mp3_dev"):
            getattr(lib, name).argtypes = [vp, C.c_int, vp, C.c_long, C.POINTER(C.c_long), vp]
        lib.mp3gpu_begin_segment.argtypes = [vp, vp]
        lib.mp3gpu_count_bits_batch.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp]
        lib.mp3gpu_set_pcm_layout.argtypes = [vp, C.c_int]
        lib.mp3gpu_set_host_delivery.argtypes = [vp, C.c_int]
        lib.mp3gpu_frame_bytes.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.mp3gpu_format_bitstream_batch.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, C.c_long, vp]
        _lib = lib
    return _lib


def _torch():
    import torch
    return torch


def stream_wave(device=0):
    """streams that fill the device exactly once in the rate loop (mp3gpu.h)"""
    n = load_library().mp3gpu_stream_wave(int(device))
    if n < 0:
        raise Mp3GpuError(f"mp3gpu_stream_wave failed ({n}): {load_library().mp3gpu_last_error().decode()}")
    return n


class Encoder:
    """One mp3gpu_ctx: S streams of identical format, state carried from call to call."""

    def __init__(self, sfreq=44100, n_ch=2, bitrate=128, max_streams=1, max_frames=2, device=0):
        self.lib = load_library()
        self.cfg = Config(sfreq, n_ch, bitrate, max_streams, max_frames, device)
        self.ctx = C.c_void_p()
        self.n_ch, self.device = n_ch, device
        rc = self.lib.mp3gpu_create(C.byref(self.cfg), C.byref(self.ctx))
        if rc != 0:
            raise Mp3GpuError(f"mp3gpu_create failed ({rc}): {self.lib.mp3gpu_last_error().decode()}")
        bpf, mb = C.c_int(), C.c_int()
        self.lib.mp3gpu_frame_geometry(self.ctx, C.byref(bpf), C.byref(mb))
        self.bits_per_frame, self.mean_bits = bpf.value, mb.value
        fb, sib = C.c_int(), C.c_int()
        self.lib.mp3gpu_frame_bytes(self.ctx, C.byref(fb), C.byref(sib))
        self.frame_bytes, self.sideinfo_bytes = fb.value, sib.value

    def close(self):
        if self.ctx:
            self.lib.mp3gpu_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise Mp3GpuError(f"{what} failed ({rc}): {self.lib.mp3gpu_last_error().decode()}")

    def reset(self, stream=None):
        """forget all per-stream state.  Without `stream`: synchronous (waits for the device, mp3gpu_reset); with a stream:
        ordered on that stream without blocking the host (mp3gpu_reset_async)."""
        if stream is None:
            self._check(self.lib.mp3gpu_reset(self.ctx), "mp3gpu_reset")
        else:
            self._check(self.lib.mp3gpu_reset_async(self.ctx, C.c_void_p(stream or 0)), "mp3gpu_reset_async")

    def set_pipeline(self, overlap):
        """False: all work of a call on the caller's stream; True: the front end of call i+1 beside the rate loop of call i (mp3gpu.h)"""
        self._check(self.lib.mp3gpu_set_pipeline(self.ctx, 1 if overlap else 0), "mp3gpu_set_pipeline")

    def set_rate_loop_segments(self, enable):
        """speculative segmentation of the rate loop for under-filled batches (default on; bit-identical results, mp3gpu.h)"""
        self._check(self.lib.mp3gpu_set_rate_loop_segments(self.ctx, 1 if enable else 0), "mp3gpu_set_rate_loop_segments")

    def rate_loop_segment_stats(self, reset=True):
        """per pass: (frames encoded, frames replayed, segments left alone, segments merged early); see mp3gpu.h"""
        out = (C.c_long * 32)()
        self._check(self.lib.mp3gpu_rate_loop_segment_stats(self.ctx, out, 1 if reset else 0), "mp3gpu_rate_loop_segment_stats")
        return [tuple(out[4 * p:4 * p + 4]) for p in range(8) if any(out[4 * p:4 * p + 4])]

    def set_psy_variant(self, name):
        """'regs' (default): FFTs as register code; 'program': the interpreted op program (A/B, both bit-identical)"""
        self._check(self.lib.mp3gpu_set_psy_variant(self.ctx, {"regs": 0, "program": 1}[name]), "mp3gpu_set_psy_variant")

    def set_front_variant(self, name):
        """arithmetic of the fused filterbank + MDCT kernel: "exact" (default), "fma" (FP64, <= 1e-12), "fp32" (<= 1e-5)"""
        self._check(self.lib.mp3gpu_set_front_variant(self.ctx, FRONT_VARIANTS[name]), "mp3gpu_set_front_variant")

    def front_variant_info(self):
        v, b = C.c_int(), C.c_int()
        self._check(self.lib.mp3gpu_get_front_variant(self.ctx, C.byref(v), C.byref(b)), "mp3gpu_get_front_variant")
        return {"name": {n: k for k, n in FRONT_VARIANTS.items()}[v.value], "bytes_per_gc": b.value}

    def reset_streams(self, first, count, stream=None):
        """restart streams [first, first+count) as new streams; the others keep their state (mp3gpu.h)"""
        self._check(self.lib.mp3gpu_reset_streams(self.ctx, int(first), int(count), C.c_void_p(stream or 0)), "mp3gpu_reset_streams")

    def set_stream_frames(self, frames, stream=None):
        """per-stream lengths in frames, counted from the last reset (None: no limits); see mp3gpu.h"""
        if frames is None:
            rc = self.lib.mp3gpu_set_stream_frames(self.ctx, 1, None, C.c_void_p(stream or 0))
        else:
            arr = (C.c_long * len(frames))(*[int(f) for f in frames])
            rc = self.lib.mp3gpu_set_stream_frames(self.ctx, len(frames), arr, C.c_void_p(stream or 0))
        self._check(rc, "mp3gpu_set_stream_frames")

    @property
    def kernel_launches(self):
        return int(self.lib.mp3gpu_kernel_launches(self.ctx))

    def profile_enable(self, on=True):
        self._check(self.lib.mp3gpu_profile_enable(self.ctx, 1 if on else 0), "mp3gpu_profile_enable")

    def profile_collect(self, reset=True):
        """-> {kernel name: (total ms, launches)} measured with CUDA events on the launching stream"""
        k = len(KERNEL_NAMES)
        ms, n = (C.c_double * k)(), (C.c_long * k)()
        self._check(self.lib.mp3gpu_profile_collect(self.ctx, ms, n, 1 if reset else 0), "mp3gpu_profile_collect")
        return {KERNEL_NAMES[i]: (ms[i], n[i]) for i in range(k)}

    def sync(self, stream=None):
        self._check(self.lib.mp3gpu_sync(self.ctx, C.c_void_p(stream or 0)), "mp3gpu_sync")

    # ---- helpers -------------------------------------------------------------------------------
    def set_host_delivery(self, pipelined):
        """host-buffer MP3 delivery: in stream order (default) or pipelined behind the next call (see mp3gpu.h)"""
        self._check(self.lib.mp3gpu_set_host_delivery(self.ctx, 1 if pipelined else 0), "mp3gpu_set_host_delivery")

    def set_pcm_layout(self, interleaved):
        """False: pcm is [S][n_ch][n] (default); True: pcm is [S][n][n_ch], the sample order of a WAV file"""
        self._check(self.lib.mp3gpu_set_pcm_layout(self.ctx, 1 if interleaved else 0), "mp3gpu_set_pcm_layout")
        self.interleaved = bool(interleaved)

    def _shape(self, pcm_shape):
        if getattr(self, "interleaved", False):
            S, n, n_ch = pcm_shape
        else:
            S, n_ch, n = pcm_shape
        assert n_ch == self.n_ch and n % 1152 == 0, (pcm_shape, self.n_ch)
        return S, n // 1152

    def _dev(self):
        return _torch().device("cuda", self.device)

    def _empty(self, shape, dtype):
        return _torch().empty(shape, dtype=dtype, device=self._dev())

    # ---- whole hot path, HOST buffers (numpy, ideally pinned via torch) ---------------------------
    def encode_frames(self, pcm, out=None, stream=None, sync=True):
        """pcm: int16 [S][n_ch][n_frames*1152] numpy array (or pinned torch CPU tensor).
        Returns dict(ix [S][gc][576] int16, gi [S][gc][20] int32, sf [S][gc][40] uint8, fo [S][frames] FO_DT)."""
        pcm_np = pcm if isinstance(pcm, np.ndarray) else pcm.numpy()
        assert pcm_np.dtype == np.int16 and pcm_np.flags["C_CONTIGUOUS"]
        S, F = self._shape(pcm_np.shape)
        gcs = F * 2 * self.n_ch
        if out is None:
            out = dict(ix=np.empty((S, gcs, 576), np.int16), gi=np.empty((S, gcs, 20), np.int32),
                       sf=np.empty((S, gcs, 40), np.uint8), fo=np.empty((S, F), FO_DT))
        p = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None
        rc = self.lib.mp3gpu_encode_frames(self.ctx, p(pcm_np), S, F, p(out.get("ix")), p(out.get("gi")), p(out.get("sf")),
                                           p(out.get("fo")), C.c_void_p(stream or 0))
        self._check(rc, "mp3gpu_encode_frames")
        if sync:
            self.sync(stream)
        return out

    # ---- whole hot path, DEVICE buffers (torch tensors) -------------------------------------------
    def encode_frames_dev(self, pcm, out=None, stream=None):
        torch = _torch()
        S, F = self._shape(tuple(pcm.shape))
        gcs = F * 2 * self.n_ch
        if out is None:
            out = dict(ix=self._empty((S, gcs, 576), torch.int16), gi=self._empty((S, gcs, 20), torch.int32),
                       sf=self._empty((S, gcs, 40), torch.uint8), fo=self._empty((S, F, FO_DT.itemsize), torch.uint8))
        rc = self.lib.mp3gpu_encode_frames_dev(self.ctx, pcm.data_ptr(), S, F, out["ix"].data_ptr(), out["gi"].data_ptr(),
                                               out["sf"].data_ptr(), out["fo"].data_ptr(), C.c_void_p(stream or 0))
        self._check(rc, "mp3gpu_encode_frames_dev")
        return out

    # ---- hot path + device bitstream formatter: PCM in, MP3 bytes out --------------------------------
    def encode_frames_mp3(self, pcm, mp3, stream=None):
        """HOST buffers. pcm int16 [S][n_ch][F*1152]; mp3 uint8 [S][stride] (absolute stream positions, see mp3gpu.h).
        Asynchronous on `stream`: call flush_mp3() (or sync()) before reading."""
        pcm_np = pcm if isinstance(pcm, np.ndarray) else pcm.numpy()
        mp3_np = mp3 if isinstance(mp3, np.ndarray) else mp3.numpy()
        assert pcm_np.dtype == np.int16 and pcm_np.flags["C_CONTIGUOUS"] and mp3_np.dtype == np.uint8 and mp3_np.flags["C_CONTIGUOUS"]
        S, F = self._shape(pcm_np.shape)
        assert mp3_np.shape[0] >= S
        rc = self.lib.mp3gpu_encode_frames_mp3(self.ctx, C.c_void_p(pcm_np.ctypes.data), S, F, C.c_void_p(mp3_np.ctypes.data),
                                               mp3_np.strides[0], C.c_void_p(stream or 0))
        self._check(rc, "mp3gpu_encode_frames_mp3")

    def encode_frames_mp3_dev(self, pcm, mp3, stream=None):
        """DEVICE tensors. mp3 may be None (format only, e.g. for timing)."""
        S, F = self._shape(tuple(pcm.shape))
        rc = self.lib.mp3gpu_encode_frames_mp3_dev(self.ctx, pcm.data_ptr(), S, F, mp3.data_ptr() if mp3 is not None else None,
                                                   mp3.stride(0) if mp3 is not None else 0, C.c_void_p(stream or 0))
        self._check(rc, "mp3gpu_encode_frames_mp3_dev")

    def flush_mp3(self, mp3, n_streams, stream=None):
        """deliver the bytes still in the window; returns the per-stream byte counts (numpy int64). Synchronises."""
        lengths = (C.c_long * n_streams)()
        if mp3 is None:
            rc = self.lib.mp3gpu_flush_mp3(self.ctx, n_streams, None, 0, lengths, C.c_void_p(stream or 0))
        elif isinstance(mp3, np.ndarray) or not mp3.is_cuda:
            a = mp3 if isinstance(mp3, np.ndarray) else mp3.numpy()
            rc = self.lib.mp3gpu_flush_mp3(self.ctx, n_streams, C.c_void_p(a.ctypes.data), a.strides[0], lengths, C.c_void_p(stream or 0))
        else:
            rc = self.lib.mp3gpu_flush_mp3_dev(self.ctx, n_streams, mp3.data_ptr(), mp3.stride(0), lengths, C.c_void_p(stream or 0))
        self._check(rc, "mp3gpu_flush_mp3")
        return np.array(lengths[:], dtype=np.int64)

    def format_bitstream_batch(self, out, mp3, stream=None):
        """III_format_bitstream batched: `out` = dict(ix, gi, sf, fo) of DEVICE tensors as the rate loop produces them."""
        S, gcs = out["ix"].shape[0], out["ix"].shape[1]
        F = gcs // (2 * self.n_ch)
        rc = self.lib.mp3gpu_format_bitstream_batch(self.ctx, out["ix"].data_ptr(), out["gi"].data_ptr(), out["sf"].data_ptr(),
                                                    out["fo"].data_ptr(), S, F, mp3.data_ptr() if mp3 is not None else None,
                                                    mp3.stride(0) if mp3 is not None else 0, C.c_void_p(stream or 0))
        self._check(rc, "mp3gpu_format_bitstream_batch")

    def begin_segment(self, stream=None):
        """segment seam: keep the signal history, empty the bit reservoir, restart the byte stream (mp3gpu.h)"""
        self._check(self.lib.mp3gpu_begin_segment(self.ctx, C.c_void_p(stream or 0)), "mp3gpu_begin_segment")

    def encode_streams(self, pcm, chunk_frames=None, n_samples=None):
        """Whole streams in one go (HOST numpy): pcm int16 [S][n_ch][n] (zero-padded to whole frames like
        get_audio(), encode.c:162-166) -> list of S `bytes`, each what the reference CLI writes for that stream
        except the one spurious byte close_bit_stream_w() appends.  n_samples: per-stream sample counts (streams of
        different lengths in one batch, mp3gpu_set_stream_frames); the rows of pcm are zero beyond them."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        S, n_ch, n = pcm.shape
        F = (n + 1151) // 1152
        if F * 1152 != n:
            pad = np.zeros((S, n_ch, F * 1152), np.int16)
            pad[:, :, :n] = pcm
            pcm = pad
        chunk = min(chunk_frames or self.cfg.max_frames, self.cfg.max_frames)
        mp3 = np.zeros((S, F * self.frame_bytes), np.uint8)
        self.reset()
        if n_samples is not None:
            self.set_stream_frames([(int(k) + 1151) // 1152 for k in n_samples])
        for f0 in range(0, F, chunk):
            f1 = min(F, f0 + chunk)
            self.encode_frames_mp3(np.ascontiguousarray(pcm[:, :, f0 * 1152:f1 * 1152]), mp3)
        lengths = self.flush_mp3(mp3, S)
        return [mp3[s, :lengths[s]].tobytes() for s in range(S)]

    # ---- stage entry points (torch device tensors in, torch device tensors out) -------------------
    def filter_subband_batch(self, pcm, stream=None):
        torch = _torch()
        S, F = self._shape(tuple(pcm.shape))
        sb = self._empty((S, F * 2 * self.n_ch, 18, 32), torch.float64)
        self._check(self.lib.mp3gpu_filter_subband_batch(self.ctx, pcm.data_ptr(), S, F, sb.data_ptr(), C.c_void_p(stream or 0)),
                    "mp3gpu_filter_subband_batch")
        return sb

    def L3psycho_anal_batch(self, pcm, stream=None):
        torch = _torch()
        S, F = self._shape(tuple(pcm.shape))
        psy = self._empty((S, F * 2 * self.n_ch, PSY_DT.itemsize), torch.uint8)
        self._check(self.lib.mp3gpu_L3psycho_anal_batch(self.ctx, pcm.data_ptr(), S, F, psy.data_ptr(), C.c_void_p(stream or 0)),
                    "mp3gpu_L3psycho_anal_batch")
        return psy

    def mdct_sub_batch(self, sb, psy, stream=None):
        torch = _torch()
        S, gcs = sb.shape[0], sb.shape[1]
        F = gcs // (2 * self.n_ch)
        xr = self._empty((S, gcs, 576), torch.float64)
        self._check(self.lib.mp3gpu_mdct_sub_batch(self.ctx, sb.data_ptr(), psy.data_ptr(), S, F, xr.data_ptr(), C.c_void_p(stream or 0)),
                    "mp3gpu_mdct_sub_batch")
        return xr

    def subband_mdct_batch(self, pcm, psy, stream=None):
        torch = _torch()
        S, F = self._shape(tuple(pcm.shape))
        xr = self._empty((S, F * 2 * self.n_ch, 576), torch.float64)
        self._check(self.lib.mp3gpu_subband_mdct_batch(self.ctx, pcm.data_ptr(), psy.data_ptr(), S, F, xr.data_ptr(),
                                                       C.c_void_p(stream or 0)), "mp3gpu_subband_mdct_batch")
        return xr

    def iteration_loop_batch(self, xr, psy, stream=None):
        torch = _torch()
        S, gcs = xr.shape[0], xr.shape[1]
        F = gcs // (2 * self.n_ch)
        out = dict(ix=self._empty((S, gcs, 576), torch.int16), gi=self._empty((S, gcs, 20), torch.int32),
                   sf=self._empty((S, gcs, 40), torch.uint8), fo=self._empty((S, F, FO_DT.itemsize), torch.uint8))
        rc = self.lib.mp3gpu_iteration_loop_batch(self.ctx, xr.data_ptr(), psy.data_ptr(), S, F, out["ix"].data_ptr(),
                                                  out["gi"].data_ptr(), out["sf"].data_ptr(), out["fo"].data_ptr(),
                                                  C.c_void_p(stream or 0))
        self._check(rc, "mp3gpu_iteration_loop_batch")
        return out

    def quantize_count_batch(self, xr_abs, q, block_type, stream=None):
        torch = _torch()
        n = xr_abs.shape[0]
        ix = self._empty((n, 576), torch.int16)
        gi = self._empty((n, 20), torch.int32)
        bits = self._empty((n,), torch.int32)
        rc = self.lib.mp3gpu_quantize_count_batch(self.ctx, xr_abs.data_ptr(), q.data_ptr(), block_type.data_ptr(), n,
                                                  ix.data_ptr(), gi.data_ptr(), bits.data_ptr(), C.c_void_p(stream or 0))
        self._check(rc, "mp3gpu_quantize_count_batch")
        return ix, gi, bits


def read_pcm_file(path, samples_per_read=2304):
    """PCM of a WAV or raw file exactly as the reference reads it on a little-endian host (musicin.c:352-368,
    encode.c:107-167): a file whose bytes 8..11 are "WAVE" has its samples at offset 0x2c (no chunk parsing), anything
    else is raw from offset 0.  Both end up little-endian: raw data is meant to be byte-swapped (encode.c:157-159) but
    SwapBytesInWords() never advances its pointer (common.c:619-628), so it only swaps the FIRST sample of each read,
    once per sample read - a net change only when a read has an odd number of samples, i.e. in a short last frame.
    That quirk is reproduced.  Returns a flat interleaved int16 array (channel count and rate come from the caller, as
    with the reference's -m / -s flags); samples_per_read = 1152 * n_ch."""
    data = open(path, "rb").read()
    wav = data[8:12] == b"WAVE"
    off = 0x2c if wav else 0
    n = max(0, (len(data) - off) // 2)
    x = np.frombuffer(data, dtype="<i2", count=n, offset=off).astype(np.int16)
    rem = n % samples_per_read
    if not wav and rem % 2 == 1:
        x[n - rem] = x[n - rem:n - rem + 1].byteswap()[0]
    return x


def encode_files(paths, sfreq=44100, n_ch=2, bitrate=128, device=0, chunk_frames=32):
    """Batch-encode PCM files of one format (WAV or raw, see read_pcm_file) to MPEG-1 Layer III byte streams on one
    GPU: the batched equivalent of running the reference CLI once per file (minus the spurious last byte of
    close_bit_stream_w, see mp3gpu.h).  Files of any mix of lengths share ONE ctx: every stream gets its own frame
    count (mp3gpu_set_stream_frames; the last frame is zero-filled as in encode.c:162-166), ends there, and is cut by
    mp3gpu_flush_mp3 exactly as BF_FlushBitstream cuts a stream encoded alone."""
    pcms = [read_pcm_file(p, 1152 * n_ch) for p in paths]
    frames = [(len(x) + 1152 * n_ch - 1) // (1152 * n_ch) for x in pcms]
    if not paths:
        return []
    F = max(frames)
    if F == 0:
        return [b""] * len(paths)
    S = len(paths)
    step = min(chunk_frames, F)
    enc = Encoder(sfreq, n_ch, bitrate, max_streams=S, max_frames=step, device=device)
    enc.set_pcm_layout(True)
    enc.set_stream_frames(frames)
    mp3 = np.zeros((S, F * enc.frame_bytes), np.uint8)
    chunk = np.zeros((S, step * 1152 * n_ch), np.int16)
    for f0 in range(0, F, step):
        f1 = min(F, f0 + step)
        lo, hi = f0 * 1152 * n_ch, f1 * 1152 * n_ch
        chunk[:] = 0
        for k, x in enumerate(pcms):
            if len(x) > lo:
                m = min(len(x), hi) - lo
                chunk[k, :m] = x[lo:lo + m]
        enc.encode_frames_mp3(np.ascontiguousarray(chunk[:, :hi - lo]).reshape(S, (f1 - f0) * 1152, n_ch), mp3)
        enc.sync()      # the staging array is reused by the next iteration
    lengths = enc.flush_mp3(mp3, S)
    out = [mp3[k, :lengths[k]].tobytes() for k in range(S)]
    enc.close()
    return out


def _count_bits_batch(self, ix, block_type, gi=None, stream=None):
    """count_bits() batched (mp3gpu_count_bits_batch): ix int16 [n][576] magnitudes, block_type int32 [n] (device
    tensors); gi int32 [n][20] in/out (address1..3 read).  Returns (gi, bits)."""
    torch = _torch()
    n = ix.shape[0]
    if gi is None:
        gi = torch.zeros((n, 20), dtype=torch.int32, device=self._dev())
    bits = self._empty((n,), torch.int32)
    rc = self.lib.mp3gpu_count_bits_batch(self.ctx, ix.data_ptr(), block_type.data_ptr(), n, gi.data_ptr(), bits.data_ptr(),
                                          C.c_void_p(stream or 0))
    self._check(rc, "mp3gpu_count_bits_batch")
    return gi, bits


Encoder.count_bits_batch = _count_bits_batch


def psy_to_numpy(psy_tensor):
    """uint8 device tensor [..., sizeof(psy_out)] -> numpy structured array PSY_DT."""
    a = psy_tensor.cpu().numpy()
    return a.view(PSY_DT).reshape(a.shape[:-1])


def psy_from_numpy(psy_np, device):
    torch = _torch()
    raw = np.ascontiguousarray(psy_np).view(np.uint8).reshape(psy_np.shape + (PSY_DT.itemsize,))
    return torch.from_numpy(raw.copy()).to(device)


def fo_to_numpy(fo_tensor):
    a = fo_tensor.cpu().numpy()
    return a.view(FO_DT).reshape(a.shape[:-1])
