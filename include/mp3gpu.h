/* mp3gpu.h — C ABI of libmp3gpu.so: the Layer III front end + rate loop of lieff/mp3-enc-bsd,
 * batched over many streams on one B200 (hand-written sm_100a CUDA, no CPU fallback).
 *
 * Drop-in boundary: the reference is a monolithic C program; what its frame loop
 * calls for the hot path are five plain C symbols (musicin.c:754,767,768,774,779).  This library
 * exports, for each of them, a BATCHED variant over n_streams x n_frames with the reference's
 * hidden function statics turned into per-stream state owned by an mp3gpu_ctx:
 *
 *   reference entry point (file:line)                        batched replacement
 *   -------------------------------------------------------  ------------------------------------
 *   window_subband  encode.c:287 / filter_subband encode.c:361  mp3gpu_filter_subband_batch
 *   mdct_sub        mdct.c:25                                 mp3gpu_mdct_sub_batch
 *   L3psycho_anal   l3psy.c:53                                mp3gpu_L3psycho_anal_batch
 *   iteration_loop  loop.c:232 (outer_loop, inner_loop,       mp3gpu_iteration_loop_batch
 *                   bin_search_StepSize, reservoir.c)
 *   quantize loop.c:1360 + count_bits loop.c:2099             mp3gpu_quantize_count_batch
 *     (calc_runlen, count1_bitcount, subdivide,
 *      bigv_tab_select/new_choose_table, bigv_bitcount)
 *   the four calls of one frame, musicin.c:751-779            mp3gpu_encode_frames[_dev]
 *
 * The single-frame legacy signatures themselves are declared in mp3gpu_legacy.h.
 *
 * Conventions: every function returns 0 on success, a negative MP3GPU_E* code otherwise (the
 * reference exit()s/abort()s instead; a library must not).  mp3gpu_last_error() gives the text.
 * All pointers are plain host or device pointers as documented; `stream` is a cudaStream_t passed
 * as void* (NULL = default stream).  A ctx is bound to one device and is not thread safe.
 *
 * Data layouts (gc = granule-channel; n_gran = 2*n_frames; per stream g = granule*n_ch + ch):
 *   pcm   int16  [n_streams][n_ch][n_frames*1152]   planar (de-interleaved, as get_audio() does)
 *   sb    double [n_streams][n_gran*n_ch][18][32]    raw filter_subband output (before mdct sign fix)
 *   xr    double [n_streams][n_gran*n_ch][576]       mdct_sub output, band major, short = [192][3]
 *   psy   mp3gpu_psy_out [n_streams][n_gran*n_ch]
 *   ix    int16  [n_streams][n_gran*n_ch][576]       quantised spectrum WITH sign (what the reference
 *                                                    has after l3bitstream.c:115-125)
 *   gi    mp3gpu_gr_info [n_streams][n_gran*n_ch]
 *   sf    uint8  [n_streams][n_gran*n_ch][40]        long: l[0..21]; short: s[sfb][window] at 3*sfb+w
 *   fo    mp3gpu_frame_out [n_streams][n_frames]
 */
#ifndef MP3GPU_H
#define MP3GPU_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MP3GPU_OK 0
#define MP3GPU_EINVAL (-1)   /* bad argument / unsupported configuration */
#define MP3GPU_ECUDA (-2)    /* CUDA runtime error (no device, launch failure, ...) */
#define MP3GPU_ENOMEM (-3)
#define MP3GPU_ESTATE (-4)   /* call sequence error (e.g. more streams than the ctx was created for) */

typedef struct mp3gpu_ctx mp3gpu_ctx;

typedef struct {
    int sfreq_hz;        /* 32000 | 44100 | 48000 (MPEG-1; the reference refuses LSF rates for Layer III) */
    int n_ch;            /* 1 | 2 (plain stereo; the reference refuses joint stereo for Layer III) */
    int bitrate_kbps;    /* MPEG-1 Layer III table value, total for all channels */
    int max_streams;     /* capacity: streams per call */
    int max_frames;      /* capacity: frames per stream per call */
    int device;          /* CUDA device ordinal */
} mp3gpu_config;

/* side info of one granule-channel (subset of gr_info, l3side.h:41-72), 20 ints */
typedef struct {
    int part2_3_length, big_values, count1, global_gain, scalefac_compress;
    int window_switching_flag, block_type, mixed_block_flag;
    int table_select[3];
    int region0_count, region1_count, preflag, scalefac_scale, count1table_select;
    int part2_length, address1, address2, address3;
} mp3gpu_gr_info;

/* L3psycho_anal outputs of one call (l3psy.h:32-34) */
typedef struct {
    double pe;
    double ratio_l[21];
    double ratio_s[36];   /* [sfb][window] */
    int block_type;
    int pad;
} mp3gpu_psy_out;

typedef struct {
    int resv_drain;        /* III_side_info_t.resvDrain */
    int main_data_begin;   /* back pointer of this frame, bytes */
    unsigned char scfsi[2][4];
} mp3gpu_frame_out;

const char *mp3gpu_last_error(void);
const char *mp3gpu_version(void);

int mp3gpu_create(const mp3gpu_config *cfg, mp3gpu_ctx **out);
void mp3gpu_destroy(mp3gpu_ctx *ctx);
/* forget all per-stream state (filterbank/MDCT history, psy history, reservoir, stream lengths): start of new streams.
 * mp3gpu_reset waits for everything in flight on the ctx's device, clears and returns when the clears are done;
 * mp3gpu_reset_async enqueues the clears on `stream` (behind the ctx's private copy streams) and does not block the host. */
int mp3gpu_reset(mp3gpu_ctx *ctx);
int mp3gpu_reset_async(mp3gpu_ctx *ctx, void *stream);
/* Streams of different lengths in one batch: stream s ends after frames[s] frames, counted from the last reset (or
 * mp3gpu_begin_segment).  Calls keep the lockstep shape [n_streams][n_frames]; the frames of a stream beyond its end are
 * ignored (their PCM is not read, no output is produced) and mp3gpu_flush_mp3 reports every stream's own length, exactly as
 * if the stream had been encoded alone.  frames == NULL removes the limits.  `frames` is a host array of n_streams longs.
 * (The reference encodes one stream of any length per process: its frame loop runs until get_audio() returns 0,
 * musicin.c:585, the last frame zero-filled, encode.c:162-166.) */
int mp3gpu_set_stream_frames(mp3gpu_ctx *ctx, int n_streams, const long *frames, void *stream);
/* restart streams [first, first + count) as new streams (zero history / psychoacoustic state / reservoir / byte window)
 * while the others keep their state; ordered on `stream`.  All streams of a ctx share the absolute frame position, so
 * this belongs at the start of a batch or right after mp3gpu_begin_segment (e.g. the first segment of a long stream, which
 * has no pre-roll and must start exactly like the reference's process does). */
int mp3gpu_reset_streams(mp3gpu_ctx *ctx, int first, int count, void *stream);
/* number of streams the rate loop keeps in flight at once (one warp per stream and frame: SMs x warps per SM).  The rate
 * loop is a persistent kernel drawing (stream, frame) work items from a queue, so any batch size keeps the device busy;
 * batches below this number run with fewer warps per SM.  Negative MP3GPU_E* on error. */
int mp3gpu_stream_wave(int device);
/* frame geometry the reference derives in musicin.c:562-572,729-746 */
int mp3gpu_frame_geometry(const mp3gpu_ctx *ctx, int *bits_per_frame, int *mean_bits);

/* Arithmetic of the fused polyphase filterbank + MDCT kernel (window_subband / filter_subband encode.c:287-409,
 * mdct_sub / mdct mdct.c:25-198).  BASELINE's north star allows <= 1e-12 relative error for an FP64 and <= 1e-5 for an FP32
 * path; Huffman counts and table selection stay bit-exact given the quantised values in every variant.
 *   EXACT (default)  every sum in the reference's order with unfused IEEE multiply / add: subband samples and xr
 *                    bit-identical to the reference, the MP3 bytes identical to the reference CLI's file.
 *   FMA              FP64 with fused multiply-add, the MDCT as time-domain-aliasing fold + 18-point DCT-IV, the matrixing
 *                    direct-form on the reference's 9-decimal coefficients: ~1e-15 relative to the granule's largest value.
 *   FMA_TC           as FMA, the matrixing as one 32 x 32 x 32 product per 32 slots on the FP64 tensor cores.
 *   FP32             FP32 throughout, the matrixing as Lee's fast 32-point DCT-III, float spectra handed to the rate loop:
 *                    ~1e-6.  The rate loop quantises what it is given, so frames are no longer byte-identical.
 * Error is measured against the largest |value| of the granule (pointwise relative error is meaningless for
 * near-zero lines).  Set before the first encode call of a batch. */
#define MP3GPU_FRONT_EXACT 0
#define MP3GPU_FRONT_FMA 1
#define MP3GPU_FRONT_FP32 2
#define MP3GPU_FRONT_FMA_TC 3   /* FMA with the 32 x 32 matrixing on the FP64 tensor cores (mma.sync m8n8k4, SASS DMMA): the A/B */
int mp3gpu_set_front_variant(mp3gpu_ctx *ctx, int variant);
int mp3gpu_get_front_variant(const mp3gpu_ctx *ctx, int *variant, int *algorithmic_bytes_per_gc);

/* Implementation of the psychoacoustic model's FFTs (subs.c:185-534) in k_psy_front — both bit-identical to the reference:
 *   REGS (default)  the split-radix dataflow as straight-line register code of a warp (two layouts, shuffles between the
 *                   real / imaginary lanes of a complex node)
 *   PROGRAM         the levelised op program interpreted in shared memory (round 1); kept as the A/B and self-check.
 * The environment variable MP3GPU_PSY_FFT=program selects PROGRAM at context creation. */
#define MP3GPU_PSY_REGS 0
#define MP3GPU_PSY_PROGRAM 1
int mp3gpu_set_psy_variant(mp3gpu_ctx *ctx, int variant);

/* Pipelining of successive calls of the mp3gpu_encode_frames* family.  SERIAL (default): all work of a call is enqueued on
 * `stream`.  OVERLAP: PCM staging, the psychoacoustic model and the filterbank + MDCT of a call run on a private
 * low-priority stream, so that they execute beside the rate loop of the PREVIOUS call (the rate loop is bound by the
 * per-stream dependency chain and leaves SM resources free whenever the batch is smaller than the device's warp slots:
 * sharded batches, segmented streams); psy results and spectra are double-buffered.  The rate loop, the bitstream kernels and
 * every copy of results stay on `stream`, which waits for the front end: results are ordered on `stream` exactly as in
 * SERIAL mode.  One difference: the device-pointer variants read `pcm` on the private stream, so the buffer must be
 * complete when the call is made (not produced by work still pending on `stream`); work enqueued on `stream` after the call
 * is ordered behind that read.  Switch modes between batches (the call synchronises the device). */
/* Speculative segmentation of the rate loop (on by default).  When a call's batch leaves at least half of the device's
 * warp slots empty, the frames of the call are cut into up to 8 segments per stream that are encoded concurrently: segment
 * 0 from the stream's state, the others from a guessed reservoir; a segment whose predecessor ended elsewhere is encoded
 * again from the true state until its state rejoins the recorded one.  The result is bit-identical to the sequential order
 * (reservoir.c:101-145) by construction — only the latency of a stream's dependency chain changes.  Longer calls (more
 * frames per call) give it more to cut.  0 switches it off (A/B, tests). */
int mp3gpu_set_rate_loop_segments(mp3gpu_ctx *ctx, int enable);
/* diagnostics since ctx creation (or the last reset = 1): for pass p = 0..7, out[4p .. 4p+3] = frames encoded, frames replayed
 * (only the reservoir bookkeeping redone), segments left alone, segments that rejoined their previous run before their end.
 * Synchronises the device. */
int mp3gpu_rate_loop_segment_stats(mp3gpu_ctx *ctx, long out[32], int reset);

#define MP3GPU_PIPELINE_SERIAL 0
#define MP3GPU_PIPELINE_OVERLAP 1
int mp3gpu_set_pipeline(mp3gpu_ctx *ctx, int mode);

/* Layout of the `pcm` argument of the mp3gpu_encode_frames* family.  PLANAR (default): [n_streams][n_ch][n_frames*1152],
 * what get_audio() leaves in buffer[2][1152] (encode.c:181-269).  INTERLEAVED: [n_streams][n_frames*1152][n_ch], the
 * sample order of a WAV / raw PCM file as read_samples() delivers it (encode.c:107-167); the channel split of
 * get_audio() (encode.c:256-269) then runs as a CUDA kernel. */
#define MP3GPU_PCM_PLANAR 0
#define MP3GPU_PCM_INTERLEAVED 1
int mp3gpu_set_pcm_layout(mp3gpu_ctx *ctx, int layout);

/* Delivery of the MP3 bytes of the HOST-buffer calls (mp3gpu_encode_frames_mp3).  INORDER (default): the device-to-host
 * copy of a call is issued on `stream`; its bytes have landed when the call's work on the stream has completed.
 * PIPELINED: the copy runs on a private stream while the kernels of the next call execute; the bytes of a call are
 * guaranteed once the work of the NEXT mp3gpu_encode_frames_mp3 call — or of mp3gpu_flush_mp3, which joins every copy
 * still in flight — has completed on its stream.  (The reference has no counterpart: it fwrite()s frame by frame,
 * formatBitstream.c:218-247.) */
#define MP3GPU_DELIVER_INORDER 0
#define MP3GPU_DELIVER_PIPELINED 1
int mp3gpu_set_host_delivery(mp3gpu_ctx *ctx, int mode);

/* ---- whole hot path: psy -> filterbank -> MDCT -> rate loop, continuing the ctx's streams ---------
 * Host variant: pcm and outputs are HOST pointers (pinned for async copies); H2D and D2H copies are
 * issued on `stream` and the call returns after they were enqueued; synchronise the stream (or call
 * mp3gpu_sync) before reading outputs.  Any output pointer may be NULL to skip that copy. */
int mp3gpu_encode_frames(mp3gpu_ctx *ctx, const int16_t *pcm, int n_streams, int n_frames,
                         int16_t *ix, mp3gpu_gr_info *gi, uint8_t *sf, mp3gpu_frame_out *fo, void *stream);
/* Device variant: same, all pointers are DEVICE pointers (e.g. torch tensors' data_ptr()). */
int mp3gpu_encode_frames_dev(mp3gpu_ctx *ctx, const int16_t *pcm, int n_streams, int n_frames,
                             int16_t *ix, mp3gpu_gr_info *gi, uint8_t *sf, mp3gpu_frame_out *fo, void *stream);
int mp3gpu_sync(mp3gpu_ctx *ctx, void *stream);

/* ---- hot path + bitstream formatting on the device: PCM in, MPEG-1 Layer III byte stream out ---------------
 * Replaces musicin.c:751-786 INCLUDING III_format_bitstream (l3bitstream.c:68-163, formatBitstream.c:53-80) for the
 * batched path: Huffman emission, scalefactors, side info, headers and the back-pointer frame assembly run as
 * CUDA kernels (the reservoir recurrence already lives in the rate-loop kernel, so nothing sequential is left
 * for the host).  mp3 is [n_streams][mp3_stride] bytes, ABSOLUTE file positions from the start of each stream:
 * successive calls continue the same streams and write further along each row (mp3_stride >= total frames x
 * frame_bytes).  Because main data of a frame may start up to 511 bytes before its header, the bytes of the last
 * few frames are only final after the NEXT call or after mp3gpu_flush_mp3(); every call delivers the bytes that
 * became final.  mp3gpu_flush_mp3() delivers the rest and reports, per stream, the length of the stream exactly as
 * the reference writes it minus the one spurious byte close_bit_stream_w() appends (common.c:968-974): the last
 * frame is cut short by (reservoir bytes) mod (main-data bytes per frame) (BF_FlushBitstream, formatBitstream.c:87-125).
 * Host variants take HOST pointers (pinned for asynchronous copies), _dev variants DEVICE pointers; `lengths` is
 * always a host array of n_streams longs (flush synchronises the stream). */
int mp3gpu_encode_frames_mp3(mp3gpu_ctx *ctx, const int16_t *pcm, int n_streams, int n_frames, uint8_t *mp3, long mp3_stride, void *stream);
int mp3gpu_encode_frames_mp3_dev(mp3gpu_ctx *ctx, const int16_t *pcm, int n_streams, int n_frames, uint8_t *mp3, long mp3_stride, void *stream);
int mp3gpu_flush_mp3(mp3gpu_ctx *ctx, int n_streams, uint8_t *mp3, long mp3_stride, long *lengths, void *stream);
int mp3gpu_flush_mp3_dev(mp3gpu_ctx *ctx, int n_streams, uint8_t *mp3, long mp3_stride, long *lengths, void *stream);
/* Segment seam: keep the signal history (filterbank, MDCT overlap, psychoacoustic state) of every stream but empty
 * its bit reservoir (reservoir.c:36 ResvSize) and restart its byte stream at position 0, so that the next frame has
 * main_data_begin = 0.  Used when one long stream is cut into segments that are encoded independently after a
 * pre-roll; the reference has no counterpart (one stream per process). */
int mp3gpu_begin_segment(mp3gpu_ctx *ctx, void *stream);
/* bytes per frame (constant: the reference never pads, musicin.c:566-581) and bytes of header + side info */
int mp3gpu_frame_bytes(const mp3gpu_ctx *ctx, int *frame_bytes, int *sideinfo_bytes);

/* ---- stage entry points (DEVICE pointers; each continues the ctx's per-stream state of that stage) */
/* window_subband + filter_subband for every slot of n_frames frames */
int mp3gpu_filter_subband_batch(mp3gpu_ctx *ctx, const int16_t *pcm, int n_streams, int n_frames, double *sb, void *stream);
/* mdct_sub: sb (raw) + block types (from psy) -> xr.  The previous granule of each channel is ctx state. */
int mp3gpu_mdct_sub_batch(mp3gpu_ctx *ctx, const double *sb, const mp3gpu_psy_out *psy, int n_streams, int n_frames,
                          double *xr, void *stream);
/* fused filter_subband + mdct_sub (the production path: subband samples never touch HBM) */
int mp3gpu_subband_mdct_batch(mp3gpu_ctx *ctx, const int16_t *pcm, const mp3gpu_psy_out *psy, int n_streams, int n_frames,
                              double *xr, void *stream);
int mp3gpu_L3psycho_anal_batch(mp3gpu_ctx *ctx, const int16_t *pcm, int n_streams, int n_frames, mp3gpu_psy_out *psy, void *stream);
int mp3gpu_iteration_loop_batch(mp3gpu_ctx *ctx, const double *xr, const mp3gpu_psy_out *psy, int n_streams, int n_frames,
                                int16_t *ix, mp3gpu_gr_info *gi, uint8_t *sf, mp3gpu_frame_out *fo, void *stream);
/* quantize() at step q[i] followed by count_bits(), n independent granules (no ctx stream state):
 * xr_abs [n][576] magnitudes, q [n], block_type [n]; writes ix [n][576] (unsigned values), gi [n]
 * (big_values,count1,count1table_select,region0/1_count,table_select,address1-3), bits [n]. */
int mp3gpu_quantize_count_batch(mp3gpu_ctx *ctx, const double *xr_abs, const int *q, const int *block_type, int n,
                                int16_t *ix, mp3gpu_gr_info *gi, int *bits, void *stream);
/* count_bits() (loop.c:2099-2113: calc_runlen, count1_bitcount, subdivide, bigv_tab_select / new_choose_table,
 * bigv_bitcount) on n given quantised granules: ix [n][576] magnitudes, block_type [n]; gi [n] is in/out: address1..3 are
 * read (subdivide() leaves them untouched when big_values == 0) and every field the reference's count_bits() sets is
 * written; bits [n] receives the return values. */
int mp3gpu_count_bits_batch(mp3gpu_ctx *ctx, const int16_t *ix, const int *block_type, int n, mp3gpu_gr_info *gi, int *bits,
                            void *stream);
/* III_format_bitstream (l3bitstream.c:68) batched: ix (signed) / gi / sf / fo as produced by the rate loop -> byte
 * stream, continuing the ctx's streams (same window semantics as mp3gpu_encode_frames_mp3_dev; mp3 may be NULL). */
int mp3gpu_format_bitstream_batch(mp3gpu_ctx *ctx, const int16_t *ix, const mp3gpu_gr_info *gi, const uint8_t *sf,
                                  const mp3gpu_frame_out *fo, int n_streams, int n_frames, uint8_t *mp3, long mp3_stride, void *stream);

/* last launch statistics: number of kernels this library launched since ctx creation */
long mp3gpu_kernel_launches(const mp3gpu_ctx *ctx);

/* per-kernel device time of the four hot kernels, measured with CUDA events on the launching stream
 * (used by bench.py for the live roofline figure).  collect() synchronises the device. */
#define MP3GPU_K_PSY_FRONT 0   /* FFTs + history-free psychoacoustics */
#define MP3GPU_K_PSY_SCAN 1    /* history-dependent psychoacoustic scan */
#define MP3GPU_K_FRONT 2       /* fused polyphase filterbank + MDCT + alias reduction */
#define MP3GPU_K_RATE_LOOP 3   /* rate loop + reservoir */
#define MP3GPU_K_BITSTREAM 4   /* header/side-info + Huffman emission kernels */
#define MP3GPU_N_KERNELS 5
int mp3gpu_profile_enable(mp3gpu_ctx *ctx, int on);
int mp3gpu_profile_collect(mp3gpu_ctx *ctx, double ms[MP3GPU_N_KERNELS], long launches[MP3GPU_N_KERNELS], int reset);

#ifdef __cplusplus
}
#endif
#endif
