// mp3gpu.cu — sm_100a kernels and the C ABI of libmp3gpu.so (include/mp3gpu.h).
//
// Kernels (all warp-centric; the per-warp algorithms live in *_core.h and are shared verbatim with
// the host-emulation test harness):
//   k_psy_front   one warp per granule-channel        FFTs + history-free psy          (psy_core.h)
//   k_psy_scan    one warp per (stream, channel)      history-dependent psy scan       (psy_core.h)
//   k_front       one warp per (stream, channel, tile) polyphase + MDCT + alias, fused (front_core.h)
//   k_rate_loop   one warp per stream                 rate loop + reservoir            (rate_loop_core.h)
//   k_mdct / k_quantize_count / k_roll_history        stage entry points and plumbing
// There is no CPU fallback: every entry point fails with MP3GPU_ECUDA if no device is usable.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stddef.h>
#include <limits.h>
#include <string.h>

#include <new>
#include <string>
#include <vector>

#include "../../include/mp3gpu.h"
#include "front_core.h"
#include "front_tile.cuh"
#include "front_fast.cuh"
#include "psy_core.h"
#include "rate_loop_core.h"
#include "tables.h"
#include "bitstream.cuh"

using namespace mp3gpu;

static_assert(sizeof(mp3gpu_gr_info) == sizeof(GrInfoOut), "gr_info layout");
static_assert(sizeof(mp3gpu_psy_out) == sizeof(PsyOut), "psy_out layout");
static_assert(sizeof(mp3gpu_frame_out) == sizeof(FrameOut), "frame_out layout");

#define HIST 1056  // PCM samples of history kept per channel: 576 (previous granule) + 480 (filterbank)



// ---------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------
#define FRONT_WARPS 4
__global__ void __launch_bounds__(FRONT_WARPS * 32)
k_front(const short *pcm, long stream_stride, long ch_stride, int n_streams, int n_ch, int n_gran, int tile,
        const PsyOut *psy, double *xr, double *sb, int do_mdct)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *window = reinterpret_cast<double *>(smem_raw);
    FrontWarpSmem *Ms = reinterpret_cast<FrontWarpSmem *>(window + 512);
    for (int i = threadIdx.x; i < 512; i += blockDim.x) window[i] = c_front.window[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const int n_tiles = (n_gran + tile - 1) / tile;
    const long wid = (long)blockIdx.x * FRONT_WARPS + warp;
    if (wid >= (long)n_streams * n_ch * n_tiles) return;
    const int t = (int)(wid % n_tiles);
    const int ch = (int)((wid / n_tiles) % n_ch);
    const long s = wid / ((long)n_tiles * n_ch);
    const int g_first = t * tile;
    const int ng = min(tile, n_gran - g_first);
    const long gc0 = (s * n_gran + g_first) * n_ch + ch;  // first granule-channel of this warp
    WarpCtx w;
    front_walk(w, c_front, window, Ms[warp], pcm + s * stream_stride + ch * ch_stride + HIST, g_first, ng,
               psy ? &psy[gc0].block_type : nullptr, (long)(sizeof(PsyOut) / sizeof(int)) * n_ch,
               xr ? xr + gc0 * 576 : nullptr, 576L * n_ch, sb ? sb + gc0 * 576 : nullptr, 576L * n_ch, do_mdct != 0);
}

// stage entry point mdct_sub: subband samples come from HBM, previous granule from ctx state
__global__ void __launch_bounds__(FRONT_WARPS * 32)
k_mdct(const double *sb, const PsyOut *psy, double *sb_prev, int n_streams, int n_ch, int n_gran, double *xr)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrontWarpSmem *Ms = reinterpret_cast<FrontWarpSmem *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long wid = (long)blockIdx.x * FRONT_WARPS + warp;
    if (wid >= (long)n_streams * n_ch) return;
    const int ch = (int)(wid % n_ch);
    const long s = wid / n_ch;
    WarpCtx w;
    PerThread<double> prev[18], cur[18];
    double *pv = sb_prev + wid * 576;
#pragma unroll
    for (int k = 0; k < 18; k++) prev[k].v = pv[k * 32 + lane];
    for (int g = 0; g < n_gran; g++) {
        const long gc = (s * n_gran + g) * n_ch + ch;
#pragma unroll
        for (int k = 0; k < 18; k++) {
            double v = sb[gc * 576 + k * 32 + lane];
            cur[k].v = ((lane & 1) && (k & 1)) ? __dmul_rn(v, -1.0) : v;  // mdct.c:57-60
        }
        mdct_store(w, c_front, Ms[warp], prev, cur, psy[gc].block_type, xr + gc * 576);
    }
#pragma unroll
    for (int k = 0; k < 18; k++) pv[k * 32 + lane] = prev[k].v;
}

#define PSYF_WARPS 8
__global__ void __launch_bounds__(PSYF_WARPS * 32, 4)   // 4 CTAs per SM (shared-memory limit): at most 64 registers
k_psy_front(PsyDev D, const short *pcm, long stream_stride, long ch_stride, int n_streams, int n_ch, int n_gran, const int *nfr, PsyMid *mid)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PsyFrontSmem *Ms = reinterpret_cast<PsyFrontSmem *>(smem_raw);
    const int warp = threadIdx.x >> 5;
    const long gc = (long)blockIdx.x * PSYF_WARPS + warp;
    if (gc >= (long)n_streams * n_gran * n_ch) return;
    const int ch = (int)(gc % n_ch);
    const int g = (int)((gc / n_ch) % n_gran);
    const long s = gc / ((long)n_ch * n_gran);
    if (nfr && g >= 2 * nfr[s]) return;          // the stream ended before this granule (mp3gpu_set_stream_frames)
    WarpCtx w{WarpCtx::Pinned()};
    psy_front(w, D, simt::pin_smem(Ms[warp]), pcm + s * stream_stride + ch * ch_stride + HIST + 576L * g, &mid[gc]);
}

// The same with the transforms in registers (fft_regs.h): 8064 B of shared memory per warp, 3 CTAs of 9 warps per SM (72
// registers; A/B gpurun_out/r2x: 8 x 3 at 80 registers 67.6 ms, 9 x 3 62.3 ms, 6 x 4 66.9, 12 x 2 70.2: the kernel is latency bound)
#ifndef PSYF2_WARPS
#define PSYF2_WARPS 9
#endif
#ifndef PSYF2_MIN_CTAS
#define PSYF2_MIN_CTAS 3
#endif
__global__ void __launch_bounds__(PSYF2_WARPS * 32, PSYF2_MIN_CTAS)
k_psy_front_regs(PsyDev D, const short *pcm, long stream_stride, long ch_stride, int n_streams, int n_ch, int n_gran, const int *nfr, PsyMid *mid)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5;
    const long gc = (long)blockIdx.x * PSYF2_WARPS + warp;
    if (gc >= (long)n_streams * n_gran * n_ch) return;
    const int ch = (int)(gc % n_ch);
    const int g = (int)((gc / n_ch) % n_gran);
    const long s = gc / ((long)n_ch * n_gran);
    if (nfr && g >= 2 * nfr[s]) return;
    WarpCtx w{WarpCtx::Pinned()};
    float *X = reinterpret_cast<float *>(smem_raw) + (size_t)warp * FFTR_X_WORDS;
    psy_front_regs(w, D, X, pcm + s * stream_stride + ch * ch_stride + HIST + 576L * g, &mid[gc]);
}
#define PSYF2_SMEM (PSYF2_WARPS * FFTR_X_WORDS * sizeof(float))

#ifndef PSYS_WARPS
#define PSYS_WARPS 4
#endif
#ifndef PSYS_MIN_CTAS
#define PSYS_MIN_CTAS 7
#endif
#ifndef PSYS_LOCKSTEP
#define PSYS_LOCKSTEP 1
#endif
__global__ void __launch_bounds__(PSYS_WARPS * 32, PSYS_MIN_CTAS)
k_psy_scan(const PsyTables *T, const PsyMid *mid, PsyChanState *states, int n_streams, int n_ch, int n_gran, const int *nfr, PsyOut *psy)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PsyScanSmem *Ms = reinterpret_cast<PsyScanSmem *>(smem_raw);
    const int warp = threadIdx.x >> 5;
    const long wid0 = (long)blockIdx.x * PSYS_WARPS + warp;
    const bool valid = wid0 < (long)n_streams * n_ch;
    const long wid = valid ? wid0 : 0;
    const int ch = (int)(wid % n_ch);
    const long s = wid / n_ch;
    WarpCtx w;
    PsyScanRegs R;
    const int n_live = !valid ? 0 : nfr ? min(n_gran, 2 * nfr[s]) : n_gran;
    if (n_live > 0) psy_scan_load(w, states[wid], R);
    for (int g = 0; g < n_gran; g++) {
#if PSYS_LOCKSTEP
        // the warps of a CTA take the granules in step: they run the same ~2000 instructions per granule, and four warps at
        // four different places of them miss the 32 KB instruction cache four times as often (ncu: stall_no_instruction 1.4)
        __syncthreads();
#else
        if (g >= n_live) break;
#endif
        if (g < n_live) {
            const long gc = (s * n_gran + g) * n_ch + ch;
            psy_scan_step(w, *T, Ms[warp], mid[gc], R, &psy[gc]);
        }
    }
    if (n_live > 0) psy_scan_store(w, states[wid], R);
}

// One CTA of 28 warps per SM: the hot tables are held once per SM and 28 x 8064 B of per-warp working set + 5.9 KB of tables
// fill the 227 KB of shared memory (3 CTAs x 8 warps left room for 24 warps only); 72 registers per thread.
#ifndef RL_WARPS
#define RL_WARPS 28
#endif
static_assert(RL_WARPS * 32 * 72 <= 65536, "the rate loop needs 72 registers per thread");
#define RL_HOT_BYTES ((sizeof(RateHot) + 15) & ~(size_t)15)
#define RL_SMEM_BYTES (RL_HOT_BYTES + RL_WARPS * sizeof(RateWarpSmem))
static_assert(sizeof(RateHot) % 16 == 0, "RateHot must be int4-copyable");
static_assert(sizeof(RateWarpSmem) % 16 == 0, "RateWarpSmem alignment");

__device__ __forceinline__ const RateHot &load_rate_hot(const RateTables *gT, unsigned char *smem_raw)
{
    const int4 *src = reinterpret_cast<const int4 *>(&gT->hot);
    int4 *dst = reinterpret_cast<int4 *>(smem_raw);
    for (int i = threadIdx.x; i < (int)(sizeof(RateHot) / 16); i += blockDim.x) dst[i] = src[i];
    __syncthreads();
    return *reinterpret_cast<const RateHot *>(smem_raw);
}

#ifndef RL_MIN_CTAS
#define RL_MIN_CTAS 1
#endif
// Rate loop, PERSISTENT: the grid is one CTA per SM (fewer when there are fewer streams than SMs) and every warp pulls work
// items from a ticket counter until none are left.  A work item is ONE FRAME of one stream; tickets are handed out frame-major
// (ticket t -> frame t / n_streams of stream t % n_streams), so the frames of a stream are taken in order and, with S streams
// in flight, S tickets apart.  The reservoir recurrence (reservoir.c:101-145) makes frame f of a stream depend on frame f - 1:
// the warp that draws (f, s) waits until sched[1 + s] — the number of frames of stream s finished in this launch — has
// reached f, then loads the stream's state (LoopStreamState + LoopLaneState, 1.1 KB), encodes the frame and publishes
// f + 1.  An awaited frame is always in the hands of a running warp (tickets are only drawn by running warps), so waiting
// cannot deadlock whatever part of the grid is resident.
// Hand-over without touching L1: the stream state is read and written with L2-coherent (.relaxed.gpu, SASS .STRONG.GPU)
// accesses, the progress word is polled the same way, and the only fence is the MEMBAR of the releasing store.  The reader
// issues its state loads after the polled value has come back (no speculation), the writer's state stores are at L2 before
// the progress word is — so no acquire fence is needed, and with it no CCTL.IVALL: an acquire per poll, or even per item,
// throws away the quantiser / noise tables that every warp of the SM keeps re-reading through L1.
// Why: with one warp owning a stream for a whole call, 10 000 streams on 4144 warp slots ran as 3 rounds for 2.41 rounds of
// work and every CTA waited for its slowest warp; below one wave the CTAs of 28 warps left most SMs empty (1250 streams =
// 45 SMs).  Frame-granular items balance to within one frame (~0.7 ms of ~60 ms), and the launch spreads min(28, S / SMs)
// warps over ALL SMs.
__device__ __forceinline__ int ld_acquire_gpu(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// polling load: served by L2, no ordering — an acquire load invalidates the SM's L1 every time it executes, and with it the
// quantiser / noise tables every other warp of the SM is reading (ncu: 66 % of the long-scoreboard stalls of the first
// persistent version were pow43[] gathers that kept missing L1 behind the pollers)
__device__ __forceinline__ int ld_relaxed_gpu(const int *p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(int *p, int v)
{
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_gpu(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- speculative segmentation of the rate loop ------------------------------------------------------------------------------
// Below one wave of streams the rate loop is bound by the dependency chain of a stream (frame f needs the reservoir frame
// f - 1 left, reservoir.c:101-145), not by the device: 1250 streams keep 9 of 28 warp slots per SM busy for the same 200 ms
// that 4144 streams take.  The chain is cut speculatively: the n_frames of a call are split into G segments per stream and
// every (stream, segment) pair — a virtual stream — gets a warp.  Segment 0 starts from the stream's true state, the others
// from a guess (the state the stream entered the call with).  A pass records the state after every frame (snapshots) and
// the state each segment ends in; in the next pass a segment whose predecessor ended in a state other than the one it
// started from runs again from the true state — and stops as soon as its state equals the snapshot of its previous run at
// the same frame, because from there on everything it would produce is what is already there.  The reservoir recurrence is
// contractive in practice (a wrong start is forgotten within a few frames), so the second pass re-encodes a few frames per
// segment; but nothing relies on that: after pass g + 1 segment g is final by induction (segment 0 is exact after pass 1,
// and a segment is re-run whenever its start differs from its predecessor's latest end), so G passes are always enough
// and the result is identical, bit for bit, to the sequential order.  Passes with nothing to do cost one near-empty launch.
struct SegArgs {
    int G, seg_frames, pass;
    LoopStreamState *fin_s;      // [2][n_streams * G]   state a segment ended in, double-buffered by pass parity
    LoopLaneState *fin_l;
    LoopStreamState *used_s;     // [n_streams * G]      state the latest run of a segment started from
    LoopLaneState *used_l;
    LoopStreamState *snap_s;     // [n_streams * n_frames] state after every frame of the latest run that reached it
    LoopLaneState *snap_l;
    int *snap_bits;              // [n_streams * n_frames][8] max_bits of the frame's granule-channels, their part2_3_length before stuffing
    unsigned long long *stats;   // [8 passes][4]: frames encoded, frames replayed, segments left alone, segments merged early
};

struct WarpLoopState {           // the rate-loop state of one stream in the registers of a warp
    LoopStreamState S;
    PerThread<int> en[4], xm[4];
};
__device__ __forceinline__ void wls_load(WarpLoopState &W, const LoopStreamState *ps, const LoopLaneState *pl, int lane)
{
    static_assert(sizeof(LoopStreamState) % 4 == 0 && sizeof(LoopStreamState) / 4 <= 32, "LoopStreamState is moved word by word");
#pragma unroll
    for (int i = 0; i < (int)(sizeof(LoopStreamState) / 4); i++) reinterpret_cast<int *>(&W.S)[i] = ld_relaxed_gpu(reinterpret_cast<const int *>(ps) + i);
#pragma unroll
    for (int i = 0; i < 4; i++) { W.en[i].v = ld_relaxed_gpu(&pl->en[i][lane]); W.xm[i].v = ld_relaxed_gpu(&pl->xm[i][lane]); }
}
__device__ __forceinline__ void wls_store(const WarpLoopState &W, LoopStreamState *ps, LoopLaneState *pl, int lane)
{
#pragma unroll
    for (int i = 0; i < 4; i++) { st_relaxed_gpu(&pl->en[i][lane], W.en[i].v); st_relaxed_gpu(&pl->xm[i][lane], W.xm[i].v); }
    if (lane < (int)(sizeof(LoopStreamState) / 4)) st_relaxed_gpu(reinterpret_cast<int *>(ps) + lane, reinterpret_cast<const int *>(&W.S)[lane]);
}
// 2: equal; 1: equal but for the reservoir level (word 0 of LoopStreamState); 0: different
__device__ __forceinline__ int wls_compare(const WarpLoopState &W, const LoopStreamState *ps, const LoopLaneState *pl, int lane)
{
    static_assert(offsetof(LoopStreamState, resv_size) == 0, "resv_size must lead LoopStreamState");
    bool ok = true, resv_ok = true;
#pragma unroll
    for (int i = 0; i < 4; i++) ok = ok && W.en[i].v == ld_relaxed_gpu(&pl->en[i][lane]) && W.xm[i].v == ld_relaxed_gpu(&pl->xm[i][lane]);
    if (lane < (int)(sizeof(LoopStreamState) / 4)) {
        const bool same = reinterpret_cast<const int *>(&W.S)[lane] == ld_relaxed_gpu(reinterpret_cast<const int *>(ps) + lane);
        if (lane == 0) resv_ok = same; else ok = ok && same;
    }
    if (!__all_sync(0xffffffffu, ok)) return 0;
    return __all_sync(0xffffffffu, resv_ok) ? 2 : 1;
}
__device__ __forceinline__ bool wls_equal(const WarpLoopState &W, const LoopStreamState *ps, const LoopLaneState *pl, int lane)
{
    return wls_compare(W, ps, pl, lane) == 2;
}

// Re-run of a frame whose incoming state differs from the previous run's only in the reservoir level: if every
// granule-channel is still given the max_bits it was given then (ResvMaxBits is flat in the reservoir level over wide
// ranges), the rate loop would retrace its steps — quantised values, side info and scalefactors stand, and only the
// reservoir bookkeeping (ResvAdjust loop.c:355, ResvFrameEnd reservoir.c:155-226: stuffing into part2_3_length, resvDrain,
// main_data_begin) is redone.  Returns false (nothing written) when some max_bits changed.
__device__ __noinline__ bool seg_replay_frame(const FrameGeom &G, LoopStreamState &S, const PsyOut *psy_f, const int *bits_prev, GrInfoOut *gi_f,
                                              FrameOut *fo_f, int lane)
{
    int resv = S.resv_size, p23[4] = {0, 0, 0, 0};
    const int mdb = resv / 8;
    for (int gr = 0; gr < 2; gr++)
        for (int ch = 0; ch < G.n_ch; ch++) {
            const int k = gr * G.n_ch + ch, i = gr * 2 + ch;
            if (resv_max_bits(G, resv, psy_f[k].pe) != bits_prev[k]) return false;
            p23[i] = bits_prev[4 + i];
            resv += G.mean_per_ch - p23[i];
        }
    LoopStreamState T = S;
    T.resv_size = resv;
    int drain = 0;
    resv_frame_end(G, T, p23, &drain);
    S.resv_size = T.resv_size;
    if (lane == 0) {
        for (int gr = 0; gr < 2; gr++)
            for (int ch = 0; ch < G.n_ch; ch++) gi_f[gr * G.n_ch + ch].part2_3_length = p23[gr * 2 + ch];
        fo_f->resv_drain = drain;
        fo_f->main_data_begin = mdb;                               // the scfsi bytes of the frame stand
    }
    return true;
}

// SEG: the instance with the segmented mode compiled in.  The plain instance is what full batches run: the segment
// bookkeeping is cold code, but this kernel is instruction-fetch sensitive (+1250 instructions cost 4 % at 10 000 streams).
template <bool SEG>
__global__ void __launch_bounds__(RL_WARPS * 32, RL_MIN_CTAS)
k_rate_loop(const RateTables *__restrict__ gT, FrameGeom G, LoopStreamState *states, LoopLaneState *lane_states, int n_streams, int n_frames,
            const int *__restrict__ nfr, int *sched, SegArgs seg, const double *__restrict__ xr, const PsyOut *__restrict__ psy, short *ix,
            GrInfoOut *gi, unsigned char *sf, FrameOut *fo)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const RateHot &H0 = load_rate_hot(gT, smem_raw);
    const int warp = threadIdx.x >> 5;
    const RateHot &H = simt::pin_smem(H0);
    RateWarpSmem &M = simt::pin_smem(reinterpret_cast<RateWarpSmem *>(smem_raw + RL_HOT_BYTES)[warp]);
    WarpCtx w{WarpCtx::Pinned()};
    const int lane = w.lane;
    const int gpf = 2 * G.n_ch;                                  // granule-channels per frame
    const int wpc = blockDim.x >> 5;
    // Three ways to hand frames to warps, one copy of the frame's code (this kernel pays for code size):
    //   segmented  a warp per (stream, segment), see above                      (batch x G <= warps of the grid)
    //   fixed      a warp per stream walks the stream's frames in order         (batch <= warps of the grid)
    //   queue      tickets, frame-major, state handed from warp to warp         (anything larger)
    const bool segmented = SEG && seg.G > 1;
    const bool fixed = !segmented && (long)n_streams <= (long)gridDim.x * wpc;
    const long total = (long)n_streams * n_frames;
    int next_f = 0, f_end = 0;
    bool started = false, clean = false;                         // clean: the state differs from the previous run's at most in the reservoir level
    long vs = 0;                                                 // segmented: virtual stream index = a ticket (segments differ widely in cost)
    WarpLoopState W;
    for (;;) {
        long s;
        int f;
        if (segmented) {
            if (!started) {                                      // first trip of a segment: draw it, find the state it starts from
                started = true;
                if (lane == 0) vs = (long)atomicAdd(reinterpret_cast<unsigned int *>(sched), 1u);
                vs = __shfl_sync(0xffffffffu, vs, 0);
                if (vs >= (long)n_streams * seg.G) break;
                s = vs / seg.G;
                const int g = (int)(vs - s * seg.G);
                const int nf = nfr ? min(n_frames, nfr[s]) : n_frames;
                const int f0 = min(g * seg.seg_frames, nf);
                f_end = min(f0 + seg.seg_frames, nf);
                const long rd = (long)((seg.pass - 1) & 1) * n_streams * seg.G, wr = (long)(seg.pass & 1) * n_streams * seg.G;
                if (g == 0 || seg.pass == 1) wls_load(W, &states[s], &lane_states[s], lane);          // true state / the guess
                else wls_load(W, &seg.fin_s[rd + vs - 1], &seg.fin_l[rd + vs - 1], lane);             // where the predecessor ended
                if (seg.pass > 1) {
                    const int cmp = wls_compare(W, &seg.used_s[vs], &seg.used_l[vs], lane);
                    if (cmp == 2) {
                        wls_load(W, &seg.fin_s[rd + vs], &seg.fin_l[rd + vs], lane);                  // nothing changed: carry the end state over
                        wls_store(W, &seg.fin_s[wr + vs], &seg.fin_l[wr + vs], lane);
                        if (lane == 0) atomicAdd(&seg.stats[4 * (seg.pass - 1) + 2], 1ull);
                        started = false; continue;
                    }
                    clean = cmp == 1;
                }
                wls_store(W, &seg.used_s[vs], &seg.used_l[vs], lane);
                if (f0 >= f_end) {                               // empty segment: pass the state through
                    wls_store(W, &seg.fin_s[wr + vs], &seg.fin_l[wr + vs], lane);
                    started = false; continue;
                }
                next_f = f0;
            }
            s = vs / seg.G;
            f = next_f++;
        } else if (fixed) {
            s = (long)blockIdx.x * wpc + warp;
            f = next_f++;
            if (s >= n_streams || f >= (nfr ? min(n_frames, nfr[s]) : n_frames)) break;
            wls_load(W, &states[s], &lane_states[s], lane);
        } else {
            long t = 0;
            if (lane == 0) t = (long)atomicAdd(reinterpret_cast<unsigned int *>(sched), 1u);
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t >= total) break;
            f = (int)(t / n_streams);
            s = t - (long)f * n_streams;
            if (nfr && f >= nfr[s]) continue;                    // the stream ended before this frame
            if (f > 0) {
                if (lane == 0) while (ld_relaxed_gpu(sched + 1 + s) < f) __nanosleep(200);
                __syncwarp();
            }
            wls_load(W, &states[s], &lane_states[s], lane);
        }
        const long g0 = (s * n_frames + f) * gpf;                // first granule-channel of the frame
        const long sn = s * n_frames + f;
        bool replayed = false;
        if (segmented && clean) {
            replayed = seg_replay_frame(G, W.S, psy + g0, seg.snap_bits + 8 * sn, gi + g0, fo + sn, lane);
            if (replayed) {                                      // the rest of the state is what the previous run left after this frame
                const int resv = W.S.resv_size;
                wls_load(W, &seg.snap_s[sn], &seg.snap_l[sn], lane);
                W.S.resv_size = resv;
            }
        }
        if (!replayed) {
            const double *xf = G.xr_f32 ? reinterpret_cast<const double *>(reinterpret_cast<const float *>(xr) + g0 * 576) : xr + g0 * 576;
            rate_loop_stream(w, H, *gT, M, G, W.S, W.en, W.xm, 1, xf, psy + g0, ix + g0 * 576, gi + g0, sf + g0 * 40, fo + sn,
                             segmented ? seg.snap_bits + 8 * sn : nullptr, segmented ? seg.snap_bits + 8 * sn + 4 : nullptr);
        }
        if (segmented) {
            const long rd = (long)((seg.pass - 1) & 1) * n_streams * seg.G, wr = (long)(seg.pass & 1) * n_streams * seg.G;
            const int cmp = seg.pass > 1 ? wls_compare(W, &seg.snap_s[sn], &seg.snap_l[sn], lane) : 0;
            if (lane == 0) atomicAdd(&seg.stats[4 * (seg.pass - 1) + (replayed ? 1 : 0)], 1ull);
            if (cmp == 2) {
                if (lane == 0) atomicAdd(&seg.stats[4 * (seg.pass - 1) + 3], 1ull);
                // merged with the previous run: the rest of the segment, and the state it ends in, stand
                wls_load(W, &seg.fin_s[rd + vs], &seg.fin_l[rd + vs], lane);
                wls_store(W, &seg.fin_s[wr + vs], &seg.fin_l[wr + vs], lane);
                started = false; clean = false; continue;
            }
            clean = cmp == 1;
            wls_store(W, &seg.snap_s[sn], &seg.snap_l[sn], lane);
            if (next_f >= f_end) {
                wls_store(W, &seg.fin_s[wr + vs], &seg.fin_l[wr + vs], lane);
                started = false; clean = false;
            }
        } else {
            wls_store(W, &states[s], &lane_states[s], lane);
            if (!fixed) {
                __syncwarp();                                    // every lane's state stores are issued before lane 0's MEMBAR + store
                if (lane == 0) st_release_gpu(sched + 1 + s, f + 1);
            }
        }
    }
}

// after the last pass: the state every stream leaves the call with is where its last segment ended
__global__ void k_seg_commit(SegArgs seg, int n_streams, LoopStreamState *states, LoopLaneState *lane_states)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long s = t >> 5;
    const int lane = (int)(t & 31);
    if (s >= n_streams) return;
    const long src = (long)(seg.pass & 1) * n_streams * seg.G + s * seg.G + seg.G - 1;
    WarpLoopState W;
    wls_load(W, &seg.fin_s[src], &seg.fin_l[src], lane);
    wls_store(W, &states[s], &lane_states[s], lane);
}

// quantize + count_bits on n independent granules.  count_only: ix[] holds magnitudes already (count_bits(), loop.c:2099)
// and gi[] carries address1..3 in (subdivide() leaves them untouched when big_values == 0, loop.c:1642-1647).
__global__ void __launch_bounds__(RL_WARPS * 32)
k_quantize_count(const RateTables *__restrict__ gT, const double *xr_abs, const int *q, const int *block_type, int n, short *ix,
                 GrInfoOut *gi, int *bits, int count_only)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const RateHot &H = load_rate_hot(gT, smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    RateWarpSmem &M = reinterpret_cast<RateWarpSmem *>(smem_raw + RL_HOT_BYTES)[warp];
    const long i = (long)blockIdx.x * RL_WARPS + warp;
    if (i >= n) return;
    WarpCtx w;
    const int bt = block_type[i];
    const bool is_short = (bt == 2), wsf = (bt != 0);
    CountResult C;
    memset(&C, 0, sizeof(C));
    int b;
    if (!count_only) {
        for (int k = 0; k < 9; k++) {
            const int s = lane + 32 * k;
            const int e0 = slot_e0(is_short, s), e1 = is_short ? e0 + 3 : e0 + 1;
            D2 x; x.x = xr_abs[i * 576 + e0]; x.y = xr_abs[i * 576 + e1];
            M.xs[s] = x;
        }
        PerThread<float> rowmax;
        refresh_pow34(w, M, rowmax);
        int qq = q[i];
        qq = qq < -256 ? -256 : (qq > 255 ? 255 : qq);
        C.kz = 9;   // nothing known about ix[] yet
        b = probe(w, H, *gT, M, is_short, wsf, qq, rowmax, C);
    } else {
        C.address1 = gi[i].address1; C.address2 = gi[i].address2; C.address3 = gi[i].address3;
        PerThread<int> nzmax, bigmax;
        int nz = -1, bg = -1;
        for (int k = 0; k < 9; k++) {
            const int s = lane + 32 * k;
            const int e0 = slot_e0(is_short, s), e1 = is_short ? e0 + 3 : e0 + 1;
            const int a = ix[i * 576 + e0], c = ix[i * 576 + e1];
            U2 v; v.x = (unsigned short)a; v.y = (unsigned short)c;
            M.ix[s] = v;
            if ((a | c) != 0) nz = s;
            if (a > 1 || c > 1) bg = s;
        }
        nzmax.v = nz; bigmax.v = bg;
        __syncwarp();
        b = count_all(w, H, M, is_short, wsf, nzmax, bigmax, C);
    }
    __syncwarp();
    if (!count_only)
        for (int k = 0; k < 9; k++) {
            const int s = lane + 32 * k;
            const int e0 = slot_e0(is_short, s), e1 = is_short ? e0 + 3 : e0 + 1;
            const U2 v = M.ix[s];
            ix[i * 576 + e0] = (short)v.x;
            ix[i * 576 + e1] = (short)v.y;
        }
    if (lane == 0) {
        GrInfoOut g;
        memset(&g, 0, sizeof(g));
        g.big_values = C.big_values; g.count1 = C.count1; g.count1table_select = C.count1table_select;
        g.region0_count = C.region0_count; g.region1_count = C.region1_count;
        g.table_select[0] = C.table_select[0]; g.table_select[1] = C.table_select[1]; g.table_select[2] = C.table_select[2];
        g.address1 = C.address1; g.address2 = C.address2; g.address3 = C.address3;
        g.block_type = bt; g.window_switching_flag = wsf;
        gi[i] = g;
        bits[i] = b;
    }
}

// keep the last HIST samples of every channel in front of the next call's samples
__global__ void k_roll_history(short *pcm, long row_stride, long n_rows, int n_new)
{
    const long row = blockIdx.x;
    if (row >= n_rows) return;
    short *p = pcm + row * row_stride;
    for (int i = threadIdx.x; i < HIST; i += blockDim.x) p[i] = p[n_new + i];
}

// get_audio()'s channel split (encode.c:256-269) for the batched path: interleaved frames [n_streams][n][n_ch]
// -> planar rows behind each row's history.  One thread per sample frame.
__global__ void k_deinterleave(const short *__restrict__ src, short *rows, long row_stride, int n_streams, int n_ch, long n)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)n_streams * n) return;
    const long s = i / n, t = i - s * n;
    if (n_ch == 2) {
        const unsigned v = reinterpret_cast<const unsigned *>(src)[i];
        rows[(2 * s) * row_stride + HIST + t] = (short)(v & 0xffffu);
        rows[(2 * s + 1) * row_stride + HIST + t] = (short)(v >> 16);
    } else {
        rows[s * row_stride + HIST + t] = src[i];
    }
}

// frames of each stream in the current call: streams advance in lockstep (frames_done) until their own end
// (mp3gpu_set_stream_frames); nfr[s] = clamp(total[s] - frames_done, 0, n_frames)
__global__ void k_call_frames(const int *__restrict__ total, long frames_done, int n_frames, int n_streams, int *nfr)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    const long left = (long)total[s] - frames_done;
    nfr[s] = left <= 0 ? 0 : (left < n_frames ? (int)left : n_frames);
}

// ---------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, const char *a = "", const char *b = "")
{
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}
#define CU(call)                                                                            \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) return fail(MP3GPU_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

struct PcmStage {
    short *buf = nullptr;  // [max_streams*n_ch][HIST + max_frames*1152]
};

struct mp3gpu_ctx {
    mp3gpu_config cfg;
    int sr;
    FrameGeom geom;
    long row;  // samples per PCM row = HIST + max_frames*1152
    // tables
    PsyTables *d_psy_tab = nullptr;
    RateTables *d_rate_tab = nullptr;
    int sm_count = 148;
    uint32_t *d_ops1024 = nullptr, *d_ops256 = nullptr;
    int *d_lv1024 = nullptr, *d_lv256 = nullptr;
    uint32_t *d_out1024 = nullptr, *d_out256 = nullptr;
    FftTwiddle *d_tw = nullptr;
    float *d_twA = nullptr;
    uint32_t *d_out_long = nullptr, *d_out_short = nullptr;
    int psy_variant = MP3GPU_PSY_REGS;
    bool fp32_v1 = false;
    PsyDev psy_dev;
    // state
    PcmStage pcm_main, pcm_fb, pcm_psy;
    PsyChanState *d_psy_state = nullptr;
    LoopStreamState *d_loop_state = nullptr;
    LoopLaneState *d_lane_state = nullptr;
    double *d_sb_prev = nullptr;
    // workspace
    PsyMid *d_mid = nullptr;
    PsyOut *d_psyout = nullptr;
    double *d_xr = nullptr;
    short *d_ix = nullptr;
    GrInfoOut *d_gi = nullptr;
    unsigned char *d_sf = nullptr;
    FrameOut *d_fo = nullptr;
    // device bitstream formatter (bitstream.cuh): tables, sliding output window, end-of-stream back pointers
    BitTables *d_bit_tab = nullptr;
    unsigned char *d_win = nullptr, *d_win_tmp = nullptr;
    int *d_next_begin = nullptr;
    int frame_bytes = 0, si_bytes = 0, tail_frames = 0;
    long wstride = 0;
    long frames_done = 0;      // frames every stream has been offered so far (streams advance in lockstep until their own end)
    long frames_done_loop = 0; // same, for the stage pipeline up to the rate loop (equal to frames_done on the mp3 entry points)
    // per-stream lengths (mp3gpu_set_stream_frames): total frames of each stream, the frames of the current call, host copy
    int *d_total = nullptr, *d_nfr = nullptr;
    std::vector<int> h_total;
    bool have_total = false;
    int front_variant = MP3GPU_FRONT_EXACT;
    // MP3GPU_PIPELINE_OVERLAP: the front end (PCM staging, psy, filterbank + MDCT) of call i + 1 runs on a private low-priority
    // stream beside the rate loop of call i; psy results, spectra and per-call frame counts are double-buffered
    int overlap = 0, ov_turn = 0;
    cudaStream_t front_stream = nullptr;
    PsyOut *d_psyout2 = nullptr;
    double *d_xr2 = nullptr;
    int *d_nfr2 = nullptr;
    cudaEvent_t ev_front_done[2] = {nullptr, nullptr}, ev_bufs_free[2] = {nullptr, nullptr}, ev_sync = nullptr;
    // speculative segmentation of the rate loop (k_rate_loop): end / start states per virtual stream, per-frame snapshots
    int segment_rate_loop = 1;
    LoopStreamState *d_seg_fin_s = nullptr, *d_seg_used_s = nullptr, *d_seg_snap_s = nullptr;
    LoopLaneState *d_seg_fin_l = nullptr, *d_seg_used_l = nullptr, *d_seg_snap_l = nullptr;
    int *d_seg_snap_bits = nullptr;
    unsigned long long *d_seg_stats = nullptr;
    size_t seg_vs_cap = 0, seg_snap_cap = 0;
    int *d_sched = nullptr;    // rate-loop work queue: [0] ticket counter, [1 + s] frames of stream s finished in this launch
    // host-PCM ingest: double-buffered dense staging filled on a private copy stream, so that the H2D copy of
    // call i+1 overlaps the kernels of call i (the caller only ever sees its own stream)
    cudaStream_t copy_stream = nullptr;
    short *h2d_stage[2] = {nullptr, nullptr};
    cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    int h2d_turn = 0;
    int pcm_layout = MP3GPU_PCM_PLANAR;
    // host-MP3 delivery (MP3GPU_DELIVER_PIPELINED): the finished bytes of a call are staged densely on the device and
    // copied to the host on a private stream while the next call's kernels run
    int deliver = MP3GPU_DELIVER_INORDER;
    cudaStream_t d2h_stream = nullptr;
    uint8_t *d2h_stage[2] = {nullptr, nullptr};
    cudaEvent_t ev_staged[2] = {nullptr, nullptr}, ev_landed[2] = {nullptr, nullptr};
    int d2h_turn = 0;
    long launches = 0;
    // per-kernel timing (mp3gpu_profile_*): events bracket every launch of the four hot kernels
    int prof_on = 0;
    std::vector<cudaEvent_t> prof_ev;     // pairs (start, stop) in use since the last collect
    std::vector<int> prof_kind;           // kernel id of each pair
    std::vector<cudaEvent_t> prof_pool;   // events recycled by collect (no cudaEventCreate inside a timed region)
    double prof_ms[MP3GPU_N_KERNELS] = {0, 0, 0, 0, 0};
    long prof_n[MP3GPU_N_KERNELS] = {0, 0, 0, 0, 0};
};

static void prof_begin(mp3gpu_ctx *c, int kind, cudaStream_t q)
{
    if (!c->prof_on) return;
    cudaEvent_t a = nullptr, b = nullptr;
    if (c->prof_pool.size() >= 2) {
        a = c->prof_pool.back(); c->prof_pool.pop_back();
        b = c->prof_pool.back(); c->prof_pool.pop_back();
    } else {
        cudaEventCreate(&a); cudaEventCreate(&b);
    }
    cudaEventRecord(a, q);
    c->prof_ev.push_back(a); c->prof_ev.push_back(b);
    c->prof_kind.push_back(kind);
}
static void prof_end(mp3gpu_ctx *c, cudaStream_t q)
{
    if (!c->prof_on) return;
    cudaEventRecord(c->prof_ev.back(), q);
}

// A ctx is bound to cfg.device: every entry point makes that device current for its duration and restores the
// caller's device afterwards (kernels, events and the __constant__ / __device__ table symbols are per device).
struct DeviceGuard {
    int prev = -1, want;
    bool ok = true;
    explicit DeviceGuard(int dev) : want(dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() { if (prev >= 0 && prev != want) cudaSetDevice(prev); }
};
#define DEV_GUARD(c)                                                                                  \
    DeviceGuard dev_guard_((c)->cfg.device);                                                           \
    if (!dev_guard_.ok) return fail(MP3GPU_ECUDA, "cudaSetDevice(ctx device) failed")

extern "C" const char *mp3gpu_last_error(void) { return g_err; }
extern "C" const char *mp3gpu_version(void) { return "mp3gpu 0.1 (sm_100a)"; }
extern "C" long mp3gpu_kernel_launches(const mp3gpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

template <class T>
static int dalloc(T **p, size_t n)
{
    cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
    if (e != cudaSuccess) return fail(MP3GPU_ENOMEM, "cudaMalloc(%s bytes): %s", std::to_string(n * sizeof(T)).c_str(), cudaGetErrorString(e));
    return 0;
}

// output map of an FFT program as the kernels read it: bin i (0..n/2) -> re | im << 16 (see FftDev::out)
static void fft_out_pairs(const FftProgram &P, std::vector<uint32_t> *o)
{
    o->assign((size_t)P.n / 2 + 1, 0u);
    auto one = [&](int i) { return (uint32_t)(FFT_SKEW((unsigned)P.out_slot[i]) | (P.out_neg[i] ? 0x8000 : 0)); };
    for (int i = 0; i <= P.n / 2; i++) (*o)[i] = one(i) | ((i > 0 ? one(P.n - i) : 0u) << 16);
}

static int upload_fft(const FftProgram &P, uint32_t **ops, int **lv, uint32_t **out, FftDev *dev)
{
    int rc;
    std::vector<uint32_t> o;
    fft_out_pairs(P, &o);
    if ((rc = dalloc(ops, P.words.size()))) return rc;
    if ((rc = dalloc(lv, P.seg_word.size()))) return rc;
    if ((rc = dalloc(out, o.size()))) return rc;
    CU(cudaMemcpy(*ops, P.words.data(), P.words.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(*lv, P.seg_word.data(), P.seg_word.size() * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(*out, o.data(), o.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    dev->words = *ops; dev->seg_word = *lv; dev->n_levels = ((int)P.seg_word.size() - 1) / FFT_CLASSES; dev->out = *out;
    return 0;
}

static int create_body(mp3gpu_ctx *c);

extern "C" int mp3gpu_create(const mp3gpu_config *cfg, mp3gpu_ctx **out)
{
    if (!cfg || !out) return fail(MP3GPU_EINVAL, "null argument");
    *out = nullptr;
    const int sr = sr_index(cfg->sfreq_hz);
    if (sr < 0) return fail(MP3GPU_EINVAL, "unsupported sampling frequency (MPEG-1 Layer III: 32000/44100/48000)");
    if (cfg->n_ch < 1 || cfg->n_ch > 2) return fail(MP3GPU_EINVAL, "n_ch must be 1 or 2");
    static const int rates[] = {32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320};
    bool ok = false;
    for (int r : rates) ok |= (r == cfg->bitrate_kbps);
    if (!ok) return fail(MP3GPU_EINVAL, "bitrate not in the MPEG-1 Layer III table");
    if (cfg->max_streams < 1 || cfg->max_frames < 1) return fail(MP3GPU_EINVAL, "max_streams/max_frames must be >= 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(MP3GPU_ECUDA, "no CUDA device: libmp3gpu has no CPU fallback");
    DeviceGuard guard(cfg->device);
    if (!guard.ok) return fail(MP3GPU_ECUDA, "cudaSetDevice(%s) failed", std::to_string(cfg->device).c_str());
    mp3gpu_ctx *c = new (std::nothrow) mp3gpu_ctx();
    if (!c) return fail(MP3GPU_ENOMEM, "out of host memory");
    c->cfg = *cfg; c->sr = sr;
    if (cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, cfg->device) != cudaSuccess || c->sm_count < 1) c->sm_count = 148;
    frame_geometry(cfg->sfreq_hz, cfg->n_ch, cfg->bitrate_kbps, &c->geom);
    c->row = HIST + (long)cfg->max_frames * 1152;
    int rc = create_body(c);
    if (rc) { mp3gpu_destroy(c); return rc; }
    *out = c;
    return 0;
}

// allocate a device object and upload its host image; every CUDA call is checked
template <class T>
static int upload(T **dst, const T *src, size_t n = 1)
{
    int rc = dalloc(dst, n);
    if (rc) return rc;
    CU(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

// tables of the register FFT (fft_regs.h): constants into c_fftr of the current device, twiddles and output maps into global memory
static int upload_fft_regs(float **twA, uint32_t **out_long, uint32_t **out_short, PsyDev *dev)
{
    int rc;
    FftRegsPlan *P = new FftRegsPlan;
    build_fft_regs_plan(P);
    cudaError_t e = cudaMemcpyToSymbol(c_fftr, &P->c, sizeof(P->c));
    rc = (e == cudaSuccess) ? 0 : fail(MP3GPU_ECUDA, "cudaMemcpyToSymbol(c_fftr): %s", cudaGetErrorString(e));
    if (!rc) rc = upload(twA, P->twA.data(), P->twA.size());
    if (!rc) rc = upload(out_long, P->out_long.data(), P->out_long.size());
    if (!rc) rc = upload(out_short, P->out_short.data(), P->out_short.size());
    delete P;
    dev->twA = *twA; dev->out_long = *out_long; dev->out_short = *out_short;
    return rc;
}

// FP32 copies of the front-end tables for the FP32 variant (front_fast.cuh)
static cudaError_t upload_front_f(const FrontTables *F)
{
    FrontTablesF *f = new FrontTablesF;
    memset(f, 0, sizeof(*f));
    for (int w = 0; w < 4; w++) for (int k = 0; k < 36; k++) f->win[w][k] = (float)F->win[w][k];
    for (int m = 0; m < 6; m++) for (int j = 0; j < 6; j++) f->dct4_s[m][j] = (float)F->dct4_s[m][j];
    for (int k = 0; k < 8; k++) { f->ca[k] = (float)F->ca[k]; f->cs[k] = (float)F->cs[k]; }
    for (int i = 0; i < 512; i++) f->window[i] = (float)F->window[i];
    for (int m = 0; m < 18; m++) for (int j = 0; j < 18; j++) f->dct4_l[m][j] = (float)F->dct4_l[m][j];
    cudaError_t e = cudaMemcpyToSymbol(c_front_f, f, sizeof(FrontTablesF));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_front_f, f, sizeof(FrontTablesF));
    delete f;
    return e;
}

static int create_body(mp3gpu_ctx *c)
{
    const mp3gpu_config *cfg = &c->cfg;
    const int sr = c->sr;
    int rc = 0;
    // ---- tables ----
    {
        FrontTables *F = new FrontTables;
        build_front_tables(F);
        cudaError_t e = cudaMemcpyToSymbol(c_front, F, sizeof(FrontTables));
        if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_front, F, sizeof(FrontTables));
        if (e == cudaSuccess) e = upload_front_f(F);
        delete F;
        if (e != cudaSuccess) return fail(MP3GPU_ECUDA, "cudaMemcpyToSymbol: %s", cudaGetErrorString(e));
        PsyTables *P = new PsyTables;
        build_psy_tables(sr, P);
        rc = upload(&c->d_psy_tab, P);
        delete P;
        if (rc) return rc;
        RateTables *R = new RateTables;
        build_rate_tables(sr, R);
        rc = upload(&c->d_rate_tab, R);
        delete R;
        if (rc) return rc;
        std::vector<FftTwiddle> tw; std::vector<int> base;
        build_fft_twiddles(&tw, &base);
        FftProgram P10, P8;
        build_fft_program(10, base, &P10);
        build_fft_program(8, base, &P8);
        if ((rc = upload(&c->d_tw, tw.data(), tw.size()))) return rc;
        if ((rc = upload_fft(P10, &c->d_ops1024, &c->d_lv1024, &c->d_out1024, &c->psy_dev.f1024))) return rc;
        if ((rc = upload_fft(P8, &c->d_ops256, &c->d_lv256, &c->d_out256, &c->psy_dev.f256))) return rc;
        c->psy_dev.T = c->d_psy_tab; c->psy_dev.tw = c->d_tw;
        if ((rc = upload_fft_regs(&c->d_twA, &c->d_out_long, &c->d_out_short, &c->psy_dev))) return rc;
        BitTables *B = new BitTables;
        build_bit_tables(sr, cfg->sfreq_hz, cfg->n_ch, cfg->bitrate_kbps, B);
        c->frame_bytes = B->frame_bytes; c->si_bytes = B->si_bytes;
        rc = upload(&c->d_bit_tab, B);
        delete B;
        if (rc) return rc;
        // main data reaches back at most 511 main-data bytes (9-bit main_data_begin)
        c->tail_frames = 511 / (c->frame_bytes - c->si_bytes) + 1;
        c->wstride = (((long)(c->tail_frames + cfg->max_frames) * c->frame_bytes + 15) / 16) * 16;
    }
    // ---- state + workspace ----
    const size_t S = cfg->max_streams, NCH = cfg->n_ch, GC = (size_t)cfg->max_frames * 2 * NCH;
    if ((rc = dalloc(&c->pcm_main.buf, S * NCH * c->row))) return rc;
    if ((rc = dalloc(&c->d_psy_state, S * NCH))) return rc;
    if ((rc = dalloc(&c->d_loop_state, S))) return rc;
    if ((rc = dalloc(&c->d_lane_state, S))) return rc;
    if ((rc = dalloc(&c->d_mid, S * GC))) return rc;
    if ((rc = dalloc(&c->d_psyout, S * GC))) return rc;
    if ((rc = dalloc(&c->d_xr, S * GC * 576))) return rc;
    if ((rc = dalloc(&c->d_ix, S * GC * 576))) return rc;
    if ((rc = dalloc(&c->d_gi, S * GC))) return rc;
    if ((rc = dalloc(&c->d_sf, S * GC * 40))) return rc;
    if ((rc = dalloc(&c->d_fo, S * (size_t)cfg->max_frames))) return rc;
    if ((rc = dalloc(&c->d_win, S * (size_t)c->wstride))) return rc;
    if ((rc = dalloc(&c->d_win_tmp, S * (size_t)c->tail_frames * c->frame_bytes))) return rc;
    if ((rc = dalloc(&c->d_next_begin, S))) return rc;
    if ((rc = dalloc(&c->d_total, S))) return rc;
    if ((rc = dalloc(&c->d_nfr, S))) return rc;
    if ((rc = dalloc(&c->d_sched, S + 1))) return rc;
    c->h_total.assign(S, INT_MAX);
    // opt in to large dynamic shared memory; all three hot kernels want the largest shared-memory carve-out
    CU(cudaFuncSetAttribute(k_psy_front, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(PSYF_WARPS * sizeof(PsyFrontSmem))));
    CU(cudaFuncSetAttribute(k_front, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(512 * 8 + FRONT_WARPS * sizeof(FrontWarpSmem))));
    CU(cudaFuncSetAttribute(k_front_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FrontTileSmem)));
    CU(cudaFuncSetAttribute(k_mdct, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(FRONT_WARPS * sizeof(FrontWarpSmem))));
    CU(cudaFuncSetAttribute(k_rate_loop<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RL_SMEM_BYTES));
    CU(cudaFuncSetAttribute(k_rate_loop<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RL_SMEM_BYTES));
    CU(cudaFuncSetAttribute(k_quantize_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RL_SMEM_BYTES));
    CU(cudaFuncSetAttribute(k_front_tile, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaFuncSetAttribute(k_front_fast<double, false, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FrontFastSmem<double, false>)));
    CU(cudaFuncSetAttribute(k_front_fast<float, true, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FrontFastSmem<float, true>)));
    CU(cudaFuncSetAttribute(k_front_fast<float, true, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FrontFastSmem<float, true>)));
    CU(cudaFuncSetAttribute(k_front_fast<double, false, double>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaFuncSetAttribute(k_front_fast<double, false, double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FrontFastSmem<double, false>)));
    CU(cudaFuncSetAttribute(k_front_fast<double, false, double, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaFuncSetAttribute(k_front_fast<float, true, float>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaFuncSetAttribute(k_front_fast<float, true, double>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaFuncSetAttribute(k_rate_loop<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaFuncSetAttribute(k_rate_loop<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaFuncSetAttribute(k_psy_front, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaFuncSetAttribute(k_psy_front_regs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PSYF2_SMEM));
    CU(cudaFuncSetAttribute(k_psy_front_regs, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaFuncSetAttribute(k_front_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FrontF32Smem)));
    CU(cudaFuncSetAttribute(k_front_f32, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    if (const char *v = getenv("MP3GPU_FRONT_FP32_V1")) c->fp32_v1 = atoi(v) != 0;
    if (const char *v = getenv("MP3GPU_PSY_FFT")) c->psy_variant = strcmp(v, "program") == 0 ? MP3GPU_PSY_PROGRAM : MP3GPU_PSY_REGS;
    return mp3gpu_reset(c);
}

extern "C" void mp3gpu_destroy(mp3gpu_ctx *c)
{
    if (!c) return;
    DeviceGuard guard(c->cfg.device);
    cudaDeviceSynchronize();          // nothing of this ctx may still be in flight when its buffers go
    for (cudaEvent_t e : c->prof_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : c->prof_pool) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) {
        if (c->ev_front_done[i]) cudaEventDestroy(c->ev_front_done[i]);
        if (c->ev_bufs_free[i]) cudaEventDestroy(c->ev_bufs_free[i]);
    }
    if (c->ev_sync) cudaEventDestroy(c->ev_sync);
    if (c->front_stream) cudaStreamDestroy(c->front_stream);
    void *ptrs[] = {c->d_seg_stats, c->d_seg_fin_s, c->d_seg_fin_l, c->d_seg_used_s, c->d_seg_used_l, c->d_seg_snap_s, c->d_seg_snap_l, c->d_seg_snap_bits,
                    c->d_psyout2, c->d_xr2, c->d_nfr2, c->d_total, c->d_nfr, c->d_sched,
                    c->d_psy_tab, c->d_rate_tab, c->d_ops1024, c->d_ops256, c->d_lv1024, c->d_lv256, c->d_out1024, c->d_out256,
                    c->d_tw, c->d_twA, c->d_out_long, c->d_out_short, c->pcm_main.buf, c->pcm_fb.buf, c->pcm_psy.buf, c->d_psy_state, c->d_loop_state, c->d_lane_state,
                    c->d_sb_prev, c->d_mid, c->d_psyout, c->d_xr, c->d_ix, c->d_gi, c->d_sf, c->d_fo,
                    c->d_bit_tab, c->d_win, c->d_win_tmp, c->d_next_begin};
    for (void *p : ptrs) if (p) cudaFree(p);
    for (int i = 0; i < 2; i++) {
        if (c->h2d_stage[i]) cudaFree(c->h2d_stage[i]);
        if (c->ev_ready[i]) cudaEventDestroy(c->ev_ready[i]);
        if (c->ev_free[i]) cudaEventDestroy(c->ev_free[i]);
    }
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (int i = 0; i < 2; i++) {
        if (c->d2h_stage[i]) cudaFree(c->d2h_stage[i]);
        if (c->ev_staged[i]) cudaEventDestroy(c->ev_staged[i]);
        if (c->ev_landed[i]) cudaEventDestroy(c->ev_landed[i]);
    }
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    delete c;
}

// every copy a pipelined delivery still has in flight must land before work enqueued on q after this point
static int join_deliveries(mp3gpu_ctx *c, cudaStream_t q)
{
    for (int i = 0; i < 2; i++)
        if (c->ev_landed[i]) CU(cudaStreamWaitEvent(q, c->ev_landed[i], 0));
    return 0;
}

// Overlap mode: state owned by the front stream (PCM rows, psy state) is about to be touched on q, or q's clears must be
// visible to the front stream's next kernels.
static int order_after_front(mp3gpu_ctx *c, cudaStream_t q)
{
    if (!c->front_stream) return 0;
    CU(cudaEventRecord(c->ev_sync, c->front_stream));
    CU(cudaStreamWaitEvent(q, c->ev_sync, 0));
    return 0;
}
static int front_after(mp3gpu_ctx *c, cudaStream_t q)
{
    if (!c->front_stream) return 0;
    CU(cudaEventRecord(c->ev_sync, q));
    CU(cudaStreamWaitEvent(c->front_stream, c->ev_sync, 0));
    return 0;
}

// forget all per-stream state, ordered on stream q behind everything the ctx has in flight on its private streams
static int reset_on(mp3gpu_ctx *c, cudaStream_t q)
{
    const size_t S = c->cfg.max_streams, NCH = c->cfg.n_ch;
    int rc = join_deliveries(c, q);
    if (rc) return rc;
    if ((rc = order_after_front(c, q))) return rc;
    for (int i = 0; i < 2; i++)
        if (c->ev_ready[i]) CU(cudaStreamWaitEvent(q, c->ev_ready[i], 0));      // host-PCM staging copies
    CU(cudaMemsetAsync(c->pcm_main.buf, 0, S * NCH * c->row * sizeof(short), q));
    if (c->pcm_fb.buf) CU(cudaMemsetAsync(c->pcm_fb.buf, 0, S * NCH * c->row * sizeof(short), q));
    if (c->pcm_psy.buf) CU(cudaMemsetAsync(c->pcm_psy.buf, 0, S * NCH * c->row * sizeof(short), q));
    if (c->d_sb_prev) CU(cudaMemsetAsync(c->d_sb_prev, 0, S * NCH * 576 * sizeof(double), q));
    CU(cudaMemsetAsync(c->d_psy_state, 0, S * NCH * sizeof(PsyChanState), q));
    CU(cudaMemsetAsync(c->d_loop_state, 0, S * sizeof(LoopStreamState), q));
    CU(cudaMemsetAsync(c->d_lane_state, 0, S * sizeof(LoopLaneState), q));
    CU(cudaMemsetAsync(c->d_win, 0, S * (size_t)c->wstride, q));
    CU(cudaMemsetAsync(c->d_next_begin, 0, S * sizeof(int), q));
    c->frames_done = 0;
    c->frames_done_loop = 0;
    c->have_total = false;
    c->h_total.assign(S, INT_MAX);
    return front_after(c, q);
}

// Synchronous reset: waits for everything in flight on the ctx's device (any stream), clears, returns when cleared.
extern "C" int mp3gpu_reset(mp3gpu_ctx *c)
{
    if (!c) return fail(MP3GPU_EINVAL, "null ctx");
    DEV_GUARD(c);
    CU(cudaDeviceSynchronize());
    int rc = reset_on(c, nullptr);
    if (rc) return rc;
    CU(cudaDeviceSynchronize());
    return 0;
}

// Stream-ordered reset: the clears are enqueued on `stream` behind the ctx's private copy streams; no host synchronisation.
extern "C" int mp3gpu_reset_async(mp3gpu_ctx *c, void *stream)
{
    if (!c) return fail(MP3GPU_EINVAL, "null ctx");
    DEV_GUARD(c);
    return reset_on(c, (cudaStream_t)stream);
}

// Restart `count` streams from stream `first` as NEW streams (zero signal history, psychoacoustic state, reservoir, byte
// window) while the others keep theirs; ordered on `stream`.  The absolute frame position of the ctx (byte offsets of the
// mp3 rows) is shared by all streams, so this is meant for the start of a batch or right after mp3gpu_begin_segment.
extern "C" int mp3gpu_reset_streams(mp3gpu_ctx *c, int first, int count, void *stream)
{
    if (!c) return fail(MP3GPU_EINVAL, "null ctx");
    if (first < 0 || count < 1 || (long)first + count > c->cfg.max_streams) return fail(MP3GPU_EINVAL, "stream range out of bounds");
    DEV_GUARD(c);
    cudaStream_t q = (cudaStream_t)stream;
    { int rco = order_after_front(c, q); if (rco) return rco; }
    const size_t NCH = c->cfg.n_ch, rows = (size_t)count * NCH, r0 = (size_t)first * NCH, pitch = (size_t)c->row * sizeof(short);
    short *bufs[3] = {c->pcm_main.buf, c->pcm_fb.buf, c->pcm_psy.buf};
    for (short *b : bufs)
        if (b) CU(cudaMemset2DAsync(b + r0 * c->row, pitch, 0, HIST * sizeof(short), rows, q));
    if (c->d_sb_prev) CU(cudaMemsetAsync(c->d_sb_prev + r0 * 576, 0, rows * 576 * sizeof(double), q));
    CU(cudaMemsetAsync(c->d_psy_state + r0, 0, rows * sizeof(PsyChanState), q));
    CU(cudaMemsetAsync(c->d_loop_state + first, 0, (size_t)count * sizeof(LoopStreamState), q));
    CU(cudaMemsetAsync(c->d_lane_state + first, 0, (size_t)count * sizeof(LoopLaneState), q));
    CU(cudaMemsetAsync(c->d_win + (size_t)first * c->wstride, 0, (size_t)count * c->wstride, q));
    CU(cudaMemsetAsync(c->d_next_begin + first, 0, (size_t)count * sizeof(int), q));
    return front_after(c, q);
}

// Per-stream lengths: stream s ends after frames[s] frames (counted from the last reset).  Calls keep the lockstep shape
// [n_streams][n_frames]; frames of a stream beyond its end are ignored (their PCM is not read, nothing is written for them)
// and mp3gpu_flush_mp3 reports each stream's own length.  frames == NULL removes the limits.  The reference encodes one
// stream of any length per process (musicin.c:585: the frame loop runs until get_audio() returns 0).
extern "C" int mp3gpu_set_stream_frames(mp3gpu_ctx *c, int n_streams, const long *frames, void *stream)
{
    if (!c) return fail(MP3GPU_EINVAL, "null ctx");
    if (n_streams < 1 || n_streams > c->cfg.max_streams) return fail(MP3GPU_EINVAL, "bad n_streams");
    DEV_GUARD(c);
    c->h_total.assign((size_t)c->cfg.max_streams, INT_MAX);
    c->have_total = frames != nullptr;
    if (!frames) return 0;
    for (int s = 0; s < n_streams; s++) {
        if (frames[s] < 0) return fail(MP3GPU_EINVAL, "negative stream length");
        c->h_total[s] = frames[s] > INT_MAX ? INT_MAX : (int)frames[s];
    }
    cudaStream_t q = (cudaStream_t)stream;
    CU(cudaMemcpyAsync(c->d_total, c->h_total.data(), c->h_total.size() * sizeof(int), cudaMemcpyHostToDevice, q));
    CU(cudaStreamSynchronize(q));     // h_total is pageable and may change before an asynchronous copy would read it
    return 0;
}

// the frames each stream contributes to a call starting at absolute frame `done` (nullptr: all of them, no limits set)
static int call_frames(mp3gpu_ctx *c, long done, int n_streams, int n_frames, cudaStream_t q, const int **nfr, int *buf = nullptr)
{
    *nfr = nullptr;
    if (!c->have_total) return 0;
    if (!buf) buf = c->d_nfr;
    k_call_frames<<<(unsigned)((n_streams + 255) / 256), 256, 0, q>>>(c->d_total, done, n_frames, n_streams, buf);
    c->launches++;
    CU(cudaGetLastError());
    *nfr = buf;
    return 0;
}

extern "C" int mp3gpu_profile_enable(mp3gpu_ctx *c, int on)
{
    if (!c) return fail(MP3GPU_EINVAL, "null ctx");
    c->prof_on = on;
    return 0;
}

// Synchronises the device, folds all recorded event pairs into per-kernel totals and returns them.
extern "C" int mp3gpu_profile_collect(mp3gpu_ctx *c, double ms[MP3GPU_N_KERNELS], long launches[MP3GPU_N_KERNELS], int reset)
{
    if (!c) return fail(MP3GPU_EINVAL, "null ctx");
    DEV_GUARD(c);
    CU(cudaDeviceSynchronize());
    for (size_t i = 0; i < c->prof_kind.size(); i++) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, c->prof_ev[2 * i], c->prof_ev[2 * i + 1]) == cudaSuccess) {
            c->prof_ms[c->prof_kind[i]] += t;
            c->prof_n[c->prof_kind[i]]++;
        }
        c->prof_pool.push_back(c->prof_ev[2 * i]); c->prof_pool.push_back(c->prof_ev[2 * i + 1]);
    }
    c->prof_ev.clear(); c->prof_kind.clear();
    for (int k = 0; k < MP3GPU_N_KERNELS; k++) {
        if (ms) ms[k] = c->prof_ms[k];
        if (launches) launches[k] = c->prof_n[k];
        if (reset) { c->prof_ms[k] = 0; c->prof_n[k] = 0; }
    }
    return 0;
}

extern "C" int mp3gpu_frame_geometry(const mp3gpu_ctx *c, int *bits_per_frame, int *mean_bits)
{
    if (!c) return fail(MP3GPU_EINVAL, "null ctx");
    if (bits_per_frame) *bits_per_frame = c->geom.bits_per_frame;
    if (mean_bits) *mean_bits = c->geom.mean_bits;
    return 0;
}

extern "C" int mp3gpu_sync(mp3gpu_ctx *c, void *stream)
{
    if (!c) return fail(MP3GPU_EINVAL, "null ctx");
    DEV_GUARD(c);
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

static int check_shape(mp3gpu_ctx *c, int n_streams, int n_frames)
{
    if (!c) return fail(MP3GPU_EINVAL, "null ctx");
    if (n_streams < 1 || n_frames < 1) return fail(MP3GPU_EINVAL, "n_streams and n_frames must be >= 1");
    if (n_streams > c->cfg.max_streams || n_frames > c->cfg.max_frames)
        return fail(MP3GPU_ESTATE, "batch exceeds the capacity the ctx was created with");
    return 0;
}

// copy the call's PCM ([n_streams*n_ch][n_frames*1152], dense) behind the history of each row
static int stage_pcm(mp3gpu_ctx *c, PcmStage &st, const int16_t *pcm, int n_streams, int n_frames, cudaMemcpyKind kind, cudaStream_t q)
{
    if (!st.buf) {
        int rc = dalloc(&st.buf, (size_t)c->cfg.max_streams * c->cfg.n_ch * c->row);
        if (rc) return rc;
        CU(cudaMemsetAsync(st.buf, 0, (size_t)c->cfg.max_streams * c->cfg.n_ch * c->row * sizeof(short), q));
    }
    const size_t w = (size_t)n_frames * 1152 * sizeof(short);
    CU(cudaMemcpy2DAsync(st.buf + HIST, c->row * sizeof(short), pcm, w, w, (size_t)n_streams * c->cfg.n_ch, kind, q));
    return 0;
}

// host PCM -> rows of the main staging buffer through the double-buffered copy stream (see mp3gpu_ctx)
static int stage_pcm_host_overlapped(mp3gpu_ctx *c, const int16_t *pcm, int n_streams, int n_frames, cudaStream_t q)
{
    const size_t cap = (size_t)c->cfg.max_streams * c->cfg.n_ch * c->cfg.max_frames * 1152;
    if (!c->copy_stream) {
        CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            int rc = dalloc(&c->h2d_stage[i], cap);
            if (rc) return rc;
            CU(cudaEventCreateWithFlags(&c->ev_ready[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&c->ev_free[i], cudaEventDisableTiming));
        }
    }
    const int t = c->h2d_turn;
    c->h2d_turn ^= 1;
    const size_t rows = (size_t)n_streams * c->cfg.n_ch, w = (size_t)n_frames * 1152 * sizeof(short);
    CU(cudaStreamWaitEvent(c->copy_stream, c->ev_free[t], 0));       // the D2D that last read this buffer is done
    CU(cudaMemcpyAsync(c->h2d_stage[t], pcm, rows * w, cudaMemcpyHostToDevice, c->copy_stream));
    CU(cudaEventRecord(c->ev_ready[t], c->copy_stream));
    CU(cudaStreamWaitEvent(q, c->ev_ready[t], 0));
    if (c->pcm_layout == MP3GPU_PCM_INTERLEAVED) {
        const long n = (long)n_frames * 1152, total = (long)n_streams * n;
        k_deinterleave<<<(unsigned)((total + 255) / 256), 256, 0, q>>>(c->h2d_stage[t], c->pcm_main.buf, c->row, n_streams, c->cfg.n_ch, n);
        c->launches++;
        CU(cudaGetLastError());
    } else {
        CU(cudaMemcpy2DAsync(c->pcm_main.buf + HIST, c->row * sizeof(short), c->h2d_stage[t], w, w, rows, cudaMemcpyDeviceToDevice, q));
    }
    CU(cudaEventRecord(c->ev_free[t], q));
    return 0;
}

static int stage_pcm_dev(mp3gpu_ctx *c, const int16_t *pcm, int n_streams, int n_frames, cudaStream_t q)
{
    if (c->pcm_layout != MP3GPU_PCM_INTERLEAVED) return stage_pcm(c, c->pcm_main, pcm, n_streams, n_frames, cudaMemcpyDeviceToDevice, q);
    const long n = (long)n_frames * 1152, total = (long)n_streams * n;
    k_deinterleave<<<(unsigned)((total + 255) / 256), 256, 0, q>>>(pcm, c->pcm_main.buf, c->row, n_streams, c->cfg.n_ch, n);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}

// diagnostics of the speculative segmentation since ctx creation: per pass p (0-based) out[4p .. 4p+3] = frames encoded,
// frames replayed (reservoir bookkeeping only), segments left alone, segments that rejoined their previous run early
extern "C" int mp3gpu_rate_loop_segment_stats(mp3gpu_ctx *c, long out[32], int reset)
{
    if (!c || !out) return fail(MP3GPU_EINVAL, "null argument");
    DEV_GUARD(c);
    for (int i = 0; i < 32; i++) out[i] = 0;
    if (!c->d_seg_stats) return 0;
    unsigned long long h[32];
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(h, c->d_seg_stats, sizeof(h), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 32; i++) out[i] = (long)h[i];
    if (reset) CU(cudaMemset(c->d_seg_stats, 0, sizeof(h)));
    return 0;
}

extern "C" int mp3gpu_set_rate_loop_segments(mp3gpu_ctx *c, int enable)
{
    if (!c) return fail(MP3GPU_EINVAL, "null ctx");
    c->segment_rate_loop = enable ? 1 : 0;
    return 0;
}

// Pipelining of successive calls of the mp3gpu_encode_frames* family (see mp3gpu.h)
extern "C" int mp3gpu_set_pipeline(mp3gpu_ctx *c, int mode)
{
    if (!c) return fail(MP3GPU_EINVAL, "null ctx");
    if (mode != MP3GPU_PIPELINE_SERIAL && mode != MP3GPU_PIPELINE_OVERLAP) return fail(MP3GPU_EINVAL, "unknown pipeline mode");
    DEV_GUARD(c);
    CU(cudaDeviceSynchronize());                       // mode switches happen between batches: nothing in flight
    if (mode == MP3GPU_PIPELINE_OVERLAP && !c->front_stream) {
        int least = 0, greatest = 0;
        CU(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CU(cudaStreamCreateWithPriority(&c->front_stream, cudaStreamNonBlocking, least));
        const size_t S = c->cfg.max_streams, GC = (size_t)c->cfg.max_frames * 2 * c->cfg.n_ch;
        int rc;
        if ((rc = dalloc(&c->d_psyout2, S * GC))) return rc;
        if ((rc = dalloc(&c->d_xr2, S * GC * 576))) return rc;
        if ((rc = dalloc(&c->d_nfr2, S))) return rc;
        for (int i = 0; i < 2; i++) {
            CU(cudaEventCreateWithFlags(&c->ev_front_done[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&c->ev_bufs_free[i], cudaEventDisableTiming));
        }
        CU(cudaEventCreateWithFlags(&c->ev_sync, cudaEventDisableTiming));
    }
    c->overlap = mode == MP3GPU_PIPELINE_OVERLAP;
    return 0;
}

extern "C" int mp3gpu_set_front_variant(mp3gpu_ctx *c, int variant)
{
    if (!c) return fail(MP3GPU_EINVAL, "null ctx");
    if (variant != MP3GPU_FRONT_EXACT && variant != MP3GPU_FRONT_FMA && variant != MP3GPU_FRONT_FP32 && variant != MP3GPU_FRONT_FMA_TC)
        return fail(MP3GPU_EINVAL, "unknown front-end variant");
    c->front_variant = variant;
    return 0;
}

extern "C" int mp3gpu_set_psy_variant(mp3gpu_ctx *c, int variant)
{
    if (!c) return fail(MP3GPU_EINVAL, "null ctx");
    if (variant != MP3GPU_PSY_REGS && variant != MP3GPU_PSY_PROGRAM) return fail(MP3GPU_EINVAL, "unknown psy variant");
    c->psy_variant = variant;
    return 0;
}

extern "C" int mp3gpu_get_front_variant(const mp3gpu_ctx *c, int *variant, int *xr_bytes_per_gc)
{
    if (!c) return fail(MP3GPU_EINVAL, "null ctx");
    if (variant) *variant = c->front_variant;
    // algorithmic bytes per granule-channel of the fused kernel: PCM + block type + xr
    if (xr_bytes_per_gc) *xr_bytes_per_gc = 1152 + 4 + (c->front_variant == MP3GPU_FRONT_FP32 ? 2304 : 4608);
    return 0;
}

extern "C" int mp3gpu_set_pcm_layout(mp3gpu_ctx *c, int layout)
{
    if (!c) return fail(MP3GPU_EINVAL, "null ctx");
    if (layout != MP3GPU_PCM_PLANAR && layout != MP3GPU_PCM_INTERLEAVED) return fail(MP3GPU_EINVAL, "unknown PCM layout");
    c->pcm_layout = layout;
    return 0;
}

extern "C" int mp3gpu_set_host_delivery(mp3gpu_ctx *c, int mode)
{
    if (!c) return fail(MP3GPU_EINVAL, "null ctx");
    if (mode != MP3GPU_DELIVER_INORDER && mode != MP3GPU_DELIVER_PIPELINED) return fail(MP3GPU_EINVAL, "unknown delivery mode");
    c->deliver = mode;
    return 0;
}

static int roll_pcm(mp3gpu_ctx *c, PcmStage &st, int n_streams, int n_frames, cudaStream_t q)
{
    k_roll_history<<<(unsigned)(n_streams * c->cfg.n_ch), 256, 0, q>>>(st.buf, c->row, (long)n_streams * c->cfg.n_ch, n_frames * 1152);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}

static int pick_tile(int n_streams, int n_ch, int n_gran)
{
    // enough warps to fill 148 SMs several times over, but tiles as long as possible (1 granule of
    // polyphase is recomputed per tile for the MDCT overlap)
    long total = (long)n_streams * n_ch * n_gran;
    long want = 148L * 16 * 4;
    long tile = total / want;
    if (tile < 1) tile = 1;
    if (tile > n_gran) tile = n_gran;
    return (int)tile;
}

static int launch_front(mp3gpu_ctx *c, const short *pcm_rows, const PsyOut *psy, int n_streams, int n_frames, double *xr, double *sb,
                        bool do_mdct, cudaStream_t q, const int *nfr = nullptr, bool f32_out = false)
{
    const int n_gran = 2 * n_frames, n_ch = c->cfg.n_ch;
    if (do_mdct && !sb) {  // production path: tiled kernel (front_tile.cuh / front_fast.cuh)
        const long n_tiles = (n_gran + FT_G - 1) / FT_G;
        const long ctas = (long)n_streams * n_ch * n_tiles;
        prof_begin(c, MP3GPU_K_FRONT, q);
        if (c->front_variant == MP3GPU_FRONT_FMA_TC) {
            k_front_fast<double, false, double, true><<<(unsigned)ctas, FT_THREADS, sizeof(FrontFastSmem<double, false>), q>>>(
                pcm_rows, c->row * n_ch, c->row, HIST, n_streams, n_ch, n_gran, nfr, psy, xr);
        } else if (c->front_variant == MP3GPU_FRONT_FMA) {
            k_front_fast<double, false, double><<<(unsigned)ctas, FT_THREADS, sizeof(FrontFastSmem<double, false>), q>>>(
                pcm_rows, c->row * n_ch, c->row, HIST, n_streams, n_ch, n_gran, nfr, psy, xr);
        } else if (c->front_variant == MP3GPU_FRONT_FP32 && f32_out) {
            if (c->fp32_v1)      // A/B: the first version of the FP32 kernel (MP3GPU_FRONT_FP32_V1=1)
                k_front_fast<float, true, float><<<(unsigned)ctas, FT_THREADS, sizeof(FrontFastSmem<float, true>), q>>>(
                    pcm_rows, c->row * n_ch, c->row, HIST, n_streams, n_ch, n_gran, nfr, psy, reinterpret_cast<float *>(xr));
            else
                k_front_f32<<<(unsigned)ctas, FT_THREADS, sizeof(FrontF32Smem), q>>>(
                    pcm_rows, c->row * n_ch, c->row, HIST, n_streams, n_ch, n_gran, nfr, psy, reinterpret_cast<float *>(xr));
        } else if (c->front_variant == MP3GPU_FRONT_FP32) {
            k_front_fast<float, true, double><<<(unsigned)ctas, FT_THREADS, sizeof(FrontFastSmem<float, true>), q>>>(
                pcm_rows, c->row * n_ch, c->row, HIST, n_streams, n_ch, n_gran, nfr, psy, xr);
        } else {
        const long grid = (FT_PERSISTENT && ctas > 2L * c->sm_count) ? 2L * c->sm_count : ctas;
        k_front_tile<<<(unsigned)grid, FT_THREADS, sizeof(FrontTileSmem), q>>>(pcm_rows, c->row * n_ch, c->row, HIST, n_streams, n_ch, n_gran, ctas, nfr, psy, xr);
        }
        prof_end(c, q);
        c->launches++;
        CU(cudaGetLastError());
        return 0;
    }
    const int tile = pick_tile(n_streams, n_ch, n_gran);
    const long n_tiles = (n_gran + tile - 1) / tile;
    const long warps = (long)n_streams * n_ch * n_tiles;
    const unsigned grid = (unsigned)((warps + FRONT_WARPS - 1) / FRONT_WARPS);
    const size_t smem = 512 * 8 + FRONT_WARPS * sizeof(FrontWarpSmem);
    prof_begin(c, MP3GPU_K_FRONT, q);
    k_front<<<grid, FRONT_WARPS * 32, smem, q>>>(pcm_rows, c->row * n_ch, c->row, n_streams, n_ch, n_gran, tile, psy, xr, sb, do_mdct ? 1 : 0);
    prof_end(c, q);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}

static int launch_psy(mp3gpu_ctx *c, const short *pcm_rows, int n_streams, int n_frames, PsyOut *psy, cudaStream_t q, const int *nfr = nullptr)
{
    const int n_gran = 2 * n_frames, n_ch = c->cfg.n_ch;
    const long gcs = (long)n_streams * n_gran * n_ch;
    prof_begin(c, MP3GPU_K_PSY_FRONT, q);
    if (c->psy_variant == MP3GPU_PSY_REGS)
        k_psy_front_regs<<<(unsigned)((gcs + PSYF2_WARPS - 1) / PSYF2_WARPS), PSYF2_WARPS * 32, PSYF2_SMEM, q>>>(
            c->psy_dev, pcm_rows, c->row * n_ch, c->row, n_streams, n_ch, n_gran, nfr, c->d_mid);
    else
        k_psy_front<<<(unsigned)((gcs + PSYF_WARPS - 1) / PSYF_WARPS), PSYF_WARPS * 32, PSYF_WARPS * sizeof(PsyFrontSmem), q>>>(
            c->psy_dev, pcm_rows, c->row * n_ch, c->row, n_streams, n_ch, n_gran, nfr, c->d_mid);
    prof_end(c, q);
    c->launches++;
    CU(cudaGetLastError());
    const long chans = (long)n_streams * n_ch;
    prof_begin(c, MP3GPU_K_PSY_SCAN, q);
    k_psy_scan<<<(unsigned)((chans + PSYS_WARPS - 1) / PSYS_WARPS), PSYS_WARPS * 32, PSYS_WARPS * sizeof(PsyScanSmem), q>>>(
        c->d_psy_tab, c->d_mid, c->d_psy_state, n_streams, n_ch, n_gran, nfr, psy);
    prof_end(c, q);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}

// warps per CTA of the persistent rate loop: the streams are spread over all SMs, at most RL_WARPS per SM
static int rate_loop_warps(const mp3gpu_ctx *c, int n_streams, unsigned *grid)
{
    int wpc = (n_streams + c->sm_count - 1) / c->sm_count;
    if (wpc > RL_WARPS) wpc = RL_WARPS;
    if (wpc < 1) wpc = 1;
    long ctas = ((long)n_streams + wpc - 1) / wpc;
    if (ctas > c->sm_count) ctas = c->sm_count;
    *grid = (unsigned)ctas;
    return wpc;
}

// Speculative segmentation (see k_rate_loop): how many segments per stream for this launch.  Only when the batch leaves
// at least half of the device's warp slots empty and the call is long enough for segments of >= 16 frames.
static int rate_loop_segments(const mp3gpu_ctx *c, int n_streams, int n_frames)
{
    if (!c->segment_rate_loop) return 1;
    const long slots = (long)c->sm_count * RL_WARPS;
    if (n_streams >= slots) return 1;                        // a full wave or more: the ticket queue keeps every warp busy
    // frames on the critical path: frames per segment + what the later passes typically re-encode behind each seam.  Only
    // while every (stream, segment) pair gets its own warp: with two rounds of segments per warp the static assignment waits
    // for the slowest pair twice and loses to the unsegmented walk (2500 heterogeneous clips, 96-frame calls: 279 ms against
    // 214 ms per step, gpurun_out/r3d)
    if (const char *v = getenv("MP3GPU_RL_SEGMENTS")) {              // A/B: force the number of segments
        const int g = atoi(v);
        return (g >= 1 && g <= 8 && n_frames / g >= 8) ? g : 1;
    }
    int best = 1;
    long best_cost = n_frames;
    for (int g = 2; g <= 8 && n_frames / g >= 16 && (long)n_streams * g <= slots; g++) {
        const long seg = (n_frames + g - 1) / g;
        const long cost = seg + 6L * (g - 1);
        if (cost * 100 < best_cost * 85) { best = g; best_cost = cost; }
    }
    return best;
}

static int launch_rate_loop(mp3gpu_ctx *c, const double *xr, const PsyOut *psy, int n_streams, int n_frames, short *ix, GrInfoOut *gi,
                            unsigned char *sf, FrameOut *fo, cudaStream_t q, const int *nfr = nullptr, bool xr_f32 = false)
{
    unsigned grid;
    FrameGeom G = c->geom;
    G.xr_f32 = xr_f32 ? 1 : 0;
    SegArgs seg;
    memset(&seg, 0, sizeof(seg));
    seg.G = rate_loop_segments(c, n_streams, n_frames);
    if (seg.G > 1) {
        const size_t vs = (size_t)n_streams * seg.G, snaps = (size_t)n_streams * n_frames;
        if (vs > c->seg_vs_cap) {
            void *old[] = {c->d_seg_fin_s, c->d_seg_fin_l, c->d_seg_used_s, c->d_seg_used_l};
            CU(cudaStreamSynchronize(q));
            for (void *p : old) if (p) cudaFree(p);
            c->d_seg_fin_s = nullptr; c->d_seg_fin_l = nullptr; c->d_seg_used_s = nullptr; c->d_seg_used_l = nullptr;
            c->seg_vs_cap = 0;
            int rc;
            if ((rc = dalloc(&c->d_seg_fin_s, 2 * vs)) || (rc = dalloc(&c->d_seg_fin_l, 2 * vs)) || (rc = dalloc(&c->d_seg_used_s, vs)) ||
                (rc = dalloc(&c->d_seg_used_l, vs))) return rc;
            c->seg_vs_cap = vs;
        }
        if (snaps > c->seg_snap_cap) {
            CU(cudaStreamSynchronize(q));
            if (c->d_seg_snap_s) cudaFree(c->d_seg_snap_s);
            if (c->d_seg_snap_l) cudaFree(c->d_seg_snap_l);
            c->d_seg_snap_s = nullptr; c->d_seg_snap_l = nullptr; c->seg_snap_cap = 0;
            int rc;
            if (c->d_seg_snap_bits) cudaFree(c->d_seg_snap_bits);
            c->d_seg_snap_bits = nullptr;
            if ((rc = dalloc(&c->d_seg_snap_s, snaps)) || (rc = dalloc(&c->d_seg_snap_l, snaps)) || (rc = dalloc(&c->d_seg_snap_bits, 8 * snaps))) return rc;
            c->seg_snap_cap = snaps;
        }
        seg.seg_frames = (n_frames + seg.G - 1) / seg.G;
        seg.fin_s = c->d_seg_fin_s; seg.fin_l = c->d_seg_fin_l; seg.used_s = c->d_seg_used_s; seg.used_l = c->d_seg_used_l;
        seg.snap_s = c->d_seg_snap_s; seg.snap_l = c->d_seg_snap_l; seg.snap_bits = c->d_seg_snap_bits;
        if (!c->d_seg_stats) {
            int rc = dalloc(&c->d_seg_stats, 32);
            if (rc) return rc;
            CU(cudaMemsetAsync(c->d_seg_stats, 0, 32 * sizeof(unsigned long long), q));
        }
        seg.stats = c->d_seg_stats;
        const long vsl = (long)vs;
        int wpc = (int)((vsl + c->sm_count - 1) / c->sm_count);
        if (wpc > RL_WARPS) wpc = RL_WARPS;
        long ctas = (vsl + wpc - 1) / wpc;
        if (ctas > c->sm_count) ctas = c->sm_count;          // more virtual streams than warps: a warp takes several, one after the other
        grid = (unsigned)ctas;
        prof_begin(c, MP3GPU_K_RATE_LOOP, q);
        for (seg.pass = 1; seg.pass <= seg.G; seg.pass++) {
            CU(cudaMemsetAsync(c->d_sched, 0, sizeof(int), q));      // the pass's ticket counter
            k_rate_loop<true><<<grid, wpc * 32, RL_HOT_BYTES + wpc * sizeof(RateWarpSmem), q>>>(c->d_rate_tab, G, c->d_loop_state, c->d_lane_state,
                                                                                                n_streams, n_frames, nfr, c->d_sched, seg, xr, psy, ix, gi, sf, fo);
            c->launches++;
        }
        seg.pass = seg.G;
        k_seg_commit<<<(unsigned)(((long)n_streams * 32 + 255) / 256), 256, 0, q>>>(seg, n_streams, c->d_loop_state, c->d_lane_state);
        c->launches++;
        prof_end(c, q);
        CU(cudaGetLastError());
        return 0;
    }
    const int wpc = rate_loop_warps(c, n_streams, &grid);
    CU(cudaMemsetAsync(c->d_sched, 0, ((size_t)n_streams + 1) * sizeof(int), q));
    prof_begin(c, MP3GPU_K_RATE_LOOP, q);
    k_rate_loop<false><<<grid, wpc * 32, RL_HOT_BYTES + wpc * sizeof(RateWarpSmem), q>>>(c->d_rate_tab, G, c->d_loop_state, c->d_lane_state,
                                                                                         n_streams, n_frames, nfr, c->d_sched, seg, xr, psy, ix, gi, sf, fo);
    prof_end(c, q);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}

static int encode_common(mp3gpu_ctx *c, const int16_t *pcm, int n_streams, int n_frames, int16_t *ix, mp3gpu_gr_info *gi, uint8_t *sf,
                         mp3gpu_frame_out *fo, void *stream, bool host, const int **nfr_out = nullptr)
{
    int rc = check_shape(c, n_streams, n_frames);
    if (rc) return rc;
    if (!pcm) return fail(MP3GPU_EINVAL, "null pcm");
    cudaStream_t q = (cudaStream_t)stream;
    const size_t gcs = (size_t)n_streams * n_frames * 2 * c->cfg.n_ch;
    // Overlap mode: everything up to the spectra runs on the private front stream `f` into buffer set b, the rate loop
    // and what follows on the caller's stream q; q waits for f, f waits for the rate loop that last read buffer set b.
    const bool ov = c->overlap != 0;
    cudaStream_t f = ov ? c->front_stream : q;
    const int b = ov ? (c->ov_turn ^= 1) : 0;
    PsyOut *psyout = b ? c->d_psyout2 : c->d_psyout;
    double *xrb = b ? c->d_xr2 : c->d_xr;
    if (ov) CU(cudaStreamWaitEvent(f, c->ev_bufs_free[b], 0));
    if (host) rc = stage_pcm_host_overlapped(c, pcm, n_streams, n_frames, f);
    else rc = stage_pcm_dev(c, pcm, n_streams, n_frames, f);
    if (rc) return rc;
    const int *nfr = nullptr;
    if ((rc = call_frames(c, c->frames_done_loop, n_streams, n_frames, f, &nfr, b ? c->d_nfr2 : c->d_nfr))) return rc;
    if (nfr_out) *nfr_out = nfr;
    // musicin.c:751-779 order: psy first (it decides block_type), then filterbank + MDCT, then the rate loop
    if ((rc = launch_psy(c, c->pcm_main.buf, n_streams, n_frames, psyout, f, nfr))) return rc;
    const bool f32 = c->front_variant == MP3GPU_FRONT_FP32;      // the FP32 front end hands float spectra to the rate loop
    if ((rc = launch_front(c, c->pcm_main.buf, psyout, n_streams, n_frames, xrb, nullptr, true, f, nfr, f32))) return rc;
    if ((rc = roll_pcm(c, c->pcm_main, n_streams, n_frames, f))) return rc;
    if (ov) {
        CU(cudaEventRecord(c->ev_front_done[b], f));
        CU(cudaStreamWaitEvent(q, c->ev_front_done[b], 0));
    }
    short *o_ix = host ? c->d_ix : (ix ? ix : c->d_ix);
    GrInfoOut *o_gi = host ? c->d_gi : (gi ? (GrInfoOut *)gi : c->d_gi);
    unsigned char *o_sf = host ? c->d_sf : (sf ? sf : c->d_sf);
    FrameOut *o_fo = host ? c->d_fo : (fo ? (FrameOut *)fo : c->d_fo);
    if ((rc = launch_rate_loop(c, xrb, psyout, n_streams, n_frames, o_ix, o_gi, o_sf, o_fo, q, nfr, f32))) return rc;
    if (ov) CU(cudaEventRecord(c->ev_bufs_free[b], q));          // re-recorded behind the formatter by the mp3 entry points
    c->frames_done_loop += n_frames;
    if (host) {
        if (ix) CU(cudaMemcpyAsync(ix, c->d_ix, gcs * 576 * sizeof(short), cudaMemcpyDeviceToHost, q));
        if (gi) CU(cudaMemcpyAsync(gi, c->d_gi, gcs * sizeof(GrInfoOut), cudaMemcpyDeviceToHost, q));
        if (sf) CU(cudaMemcpyAsync(sf, c->d_sf, gcs * 40, cudaMemcpyDeviceToHost, q));
        if (fo) CU(cudaMemcpyAsync(fo, c->d_fo, (size_t)n_streams * n_frames * sizeof(FrameOut), cudaMemcpyDeviceToHost, q));
    }
    return 0;
}

// ---- device bitstream formatter -----------------------------------------------------------------------
// Formats the call's frames into the sliding window, copies the bytes that are now final to the caller's
// buffer (absolute file positions), then slides the window.  After the call window byte 0 corresponds to
// absolute byte (frames_done - tail_frames) * FB.
static int format_common(mp3gpu_ctx *c, const short *ix, const GrInfoOut *gi, const unsigned char *sf, const FrameOut *fo, int n_streams,
                         int n_frames, uint8_t *mp3, long stride, bool host, cudaStream_t q, const int *nfr)
{
    const long FB = c->frame_bytes, T = c->tail_frames;
    if (mp3 && stride < (c->frames_done + n_frames) * FB) return fail(MP3GPU_EINVAL, "mp3 stride too small for the frames encoded so far");
    BitsGeom G;
    G.n_streams = n_streams; G.n_frames = n_frames; G.n_ch = c->cfg.n_ch;
    G.frame_bytes = c->frame_bytes; G.si_bytes = c->si_bytes;
    G.frame0 = c->frames_done; G.origin = (c->frames_done - T) * FB; G.wstride = c->wstride;
    CU(cudaMemset2DAsync(c->d_win + T * FB, c->wstride, 0, (size_t)n_frames * FB, n_streams, q));
    prof_begin(c, MP3GPU_K_BITSTREAM, q);
    const long frames = (long)n_streams * n_frames;
    k_bits_headers<<<(unsigned)((frames + 127) / 128), 128, 0, q>>>(c->d_bit_tab, G, nfr, gi, fo, c->d_win, c->d_next_begin);
    const long gcs = frames * 2 * c->cfg.n_ch;
    k_bits_emit<<<(unsigned)((gcs + BITS_WARPS - 1) / BITS_WARPS), BITS_WARPS * 32, 0, q>>>(c->d_bit_tab, G, nfr, ix, gi, sf, fo, c->d_win);
    prof_end(c, q);
    c->launches += 2;
    CU(cudaGetLastError());
    const cudaMemcpyKind kind = host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (mp3) {  // window bytes [0, n_frames*FB) are final; clip what lies before the start of the stream
        const long skip = G.origin < 0 ? -G.origin : 0, width = (long)n_frames * FB - skip;
        if (width > 0 && host && c->deliver == MP3GPU_DELIVER_PIPELINED) {
            if (!c->d2h_stream) {
                CU(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
                for (int i = 0; i < 2; i++) {
                    int rc = dalloc(&c->d2h_stage[i], (size_t)c->cfg.max_streams * c->cfg.max_frames * FB);
                    if (rc) return rc;
                    CU(cudaEventCreateWithFlags(&c->ev_staged[i], cudaEventDisableTiming));
                    CU(cudaEventCreateWithFlags(&c->ev_landed[i], cudaEventDisableTiming));
                }
            }
            const int t = c->d2h_turn;
            c->d2h_turn ^= 1;
            // q waits for BOTH copies still in flight: the one that last used this staging buffer (call i-2) and the one of
            // call i-1 — the contract in mp3gpu.h: the bytes of a call have landed once the next call's work on its stream has
            // completed.  The wait sits behind this call's kernels, so the copy of call i-1 still overlaps them.
            { int rcj = join_deliveries(c, q); if (rcj) return rcj; }
            CU(cudaMemcpy2DAsync(c->d2h_stage[t], (size_t)width, c->d_win + skip, (size_t)c->wstride, (size_t)width, n_streams,
                                 cudaMemcpyDeviceToDevice, q));
            CU(cudaEventRecord(c->ev_staged[t], q));
            CU(cudaStreamWaitEvent(c->d2h_stream, c->ev_staged[t], 0));
            CU(cudaMemcpy2DAsync(mp3 + G.origin + skip, (size_t)stride, c->d2h_stage[t], (size_t)width, (size_t)width, n_streams,
                                 cudaMemcpyDeviceToHost, c->d2h_stream));
            CU(cudaEventRecord(c->ev_landed[t], c->d2h_stream));
        } else if (width > 0)
            CU(cudaMemcpy2DAsync(mp3 + G.origin + skip, (size_t)stride, c->d_win + skip, (size_t)c->wstride, (size_t)width, n_streams, kind, q));
    }
    // slide: the last `tail` frames become the first ones (through a temporary: the ranges may overlap)
    CU(cudaMemcpy2DAsync(c->d_win_tmp, (size_t)(T * FB), c->d_win + (long)n_frames * FB, (size_t)c->wstride, (size_t)(T * FB), n_streams,
                         cudaMemcpyDeviceToDevice, q));
    CU(cudaMemcpy2DAsync(c->d_win, (size_t)c->wstride, c->d_win_tmp, (size_t)(T * FB), (size_t)(T * FB), n_streams, cudaMemcpyDeviceToDevice, q));
    c->frames_done += n_frames;
    return 0;
}

static int flush_common(mp3gpu_ctx *c, int n_streams, uint8_t *mp3, long stride, long *lengths, bool host, cudaStream_t q)
{
    if (!c) return fail(MP3GPU_EINVAL, "null ctx");
    if (n_streams < 1 || n_streams > c->cfg.max_streams) return fail(MP3GPU_EINVAL, "bad n_streams");
    DEV_GUARD(c);
    const long FB = c->frame_bytes, T = c->tail_frames;
    const long origin = (c->frames_done - T) * FB, skip = origin < 0 ? -origin : 0, width = T * FB - skip;
    if (mp3 && stride < c->frames_done * FB) return fail(MP3GPU_EINVAL, "mp3 stride too small");
    if (host) { int rc = join_deliveries(c, q); if (rc) return rc; }
    if (mp3 && width > 0)
        CU(cudaMemcpy2DAsync(mp3 + origin + skip, (size_t)stride, c->d_win + skip, (size_t)c->wstride, (size_t)width, n_streams,
                             host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, q));
    if (lengths) {
        std::vector<int> nb(n_streams);
        CU(cudaMemcpyAsync(nb.data(), c->d_next_begin, n_streams * sizeof(int), cudaMemcpyDeviceToHost, q));
        CU(cudaStreamSynchronize(q));
        // BF_FlushBitstream (formatBitstream.c:87-125) zero-fills main data for every header still queued, i.e. whole
        // frame capacities: the write position keeps its offset inside the frame, so the last frame ends up short by
        // (bytes still in the reservoir) mod (main-data bytes per frame)
        for (int s = 0; s < n_streams; s++) {
            const long fr = c->h_total[s] < c->frames_done ? (long)c->h_total[s] : c->frames_done;     // the stream's own frame count
            lengths[s] = fr * FB - nb[s] % (FB - c->si_bytes);
        }
    }
    return 0;
}

// Segment seam (single long stream cut over several GPUs): keep filterbank / MDCT / psy history, empty the bit
// reservoir of every stream and restart the byte stream, so the next frame has main_data_begin = 0.
extern "C" int mp3gpu_begin_segment(mp3gpu_ctx *c, void *stream)
{
    if (!c) return fail(MP3GPU_EINVAL, "null ctx");
    DEV_GUARD(c);
    cudaStream_t q = (cudaStream_t)stream;
    { int rcj = join_deliveries(c, q); if (rcj) return rcj; }     // the window is about to be cleared: pipelined copies of it must have landed
    static_assert(offsetof(LoopStreamState, resv_size) == 0, "resv_size must lead LoopStreamState");
    CU(cudaMemset2DAsync(c->d_loop_state, sizeof(LoopStreamState), 0, sizeof(int), (size_t)c->cfg.max_streams, q));
    CU(cudaMemsetAsync(c->d_win, 0, (size_t)c->cfg.max_streams * c->wstride, q));
    CU(cudaMemsetAsync(c->d_next_begin, 0, (size_t)c->cfg.max_streams * sizeof(int), q));
    c->frames_done = 0;
    c->frames_done_loop = 0;          // stream lengths (mp3gpu_set_stream_frames) count from the segment start
    return 0;
}

// Streams that fill the GPU exactly once in the rate loop (one warp per stream): SMs x resident CTAs x warps per CTA.
// Batches that are a multiple of this have no partially filled last wave.
extern "C" int mp3gpu_stream_wave(int device)
{
    int sms = 0, ctas = 0;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(MP3GPU_ECUDA, "cudaSetDevice failed");
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return fail(MP3GPU_ECUDA, "no device attribute");
    cudaFuncSetAttribute(k_rate_loop<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RL_SMEM_BYTES);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, k_rate_loop<false>, RL_WARPS * 32, RL_SMEM_BYTES) != cudaSuccess || ctas < 1)
        return fail(MP3GPU_ECUDA, "occupancy query failed");
    return sms * ctas * RL_WARPS;
}

extern "C" int mp3gpu_frame_bytes(const mp3gpu_ctx *c, int *frame_bytes, int *sideinfo_bytes)
{
    if (!c) return fail(MP3GPU_EINVAL, "null ctx");
    if (frame_bytes) *frame_bytes = c->frame_bytes;
    if (sideinfo_bytes) *sideinfo_bytes = c->si_bytes;
    return 0;
}

extern "C" int mp3gpu_format_bitstream_batch(mp3gpu_ctx *c, const int16_t *ix, const mp3gpu_gr_info *gi, const uint8_t *sf,
                                             const mp3gpu_frame_out *fo, int n_streams, int n_frames, uint8_t *mp3, long mp3_stride, void *stream)
{
    int rc = check_shape(c, n_streams, n_frames);
    if (rc) return rc;
    if (!ix || !gi || !sf || !fo) return fail(MP3GPU_EINVAL, "null pointer");
    DEV_GUARD(c);
    const int *nfr = nullptr;
    if ((rc = call_frames(c, c->frames_done, n_streams, n_frames, (cudaStream_t)stream, &nfr))) return rc;
    return format_common(c, ix, (const GrInfoOut *)gi, sf, (const FrameOut *)fo, n_streams, n_frames, mp3, mp3_stride, false, (cudaStream_t)stream, nfr);
}

static int encode_mp3_common(mp3gpu_ctx *c, const int16_t *pcm, int n_streams, int n_frames, uint8_t *mp3, long stride, void *stream, bool host)
{
    int rc = check_shape(c, n_streams, n_frames);
    if (rc) return rc;
    // validate before any stream state advances
    if (mp3 && stride < (c->frames_done + n_frames) * (long)c->frame_bytes)
        return fail(MP3GPU_EINVAL, "mp3 stride too small for the frames encoded so far");
    DEV_GUARD(c);
    if (c->frames_done_loop != c->frames_done) return fail(MP3GPU_ESTATE, "mp3 and non-mp3 encode calls mixed on one ctx without a reset");
    const int *nfr = nullptr;
    rc = encode_common(c, pcm, n_streams, n_frames, nullptr, nullptr, nullptr, nullptr, stream, host, &nfr);
    if (rc) return rc;
    rc = format_common(c, c->d_ix, c->d_gi, c->d_sf, c->d_fo, n_streams, n_frames, mp3, stride, host, (cudaStream_t)stream, nfr);
    if (!rc && c->overlap) CU(cudaEventRecord(c->ev_bufs_free[c->ov_turn], (cudaStream_t)stream));
    return rc;
}

extern "C" int mp3gpu_encode_frames_mp3(mp3gpu_ctx *c, const int16_t *pcm, int n_streams, int n_frames, uint8_t *mp3, long mp3_stride, void *stream)
{
    return encode_mp3_common(c, pcm, n_streams, n_frames, mp3, mp3_stride, stream, true);
}

extern "C" int mp3gpu_encode_frames_mp3_dev(mp3gpu_ctx *c, const int16_t *pcm, int n_streams, int n_frames, uint8_t *mp3, long mp3_stride,
                                            void *stream)
{
    return encode_mp3_common(c, pcm, n_streams, n_frames, mp3, mp3_stride, stream, false);
}

extern "C" int mp3gpu_flush_mp3(mp3gpu_ctx *c, int n_streams, uint8_t *mp3, long mp3_stride, long *lengths, void *stream)
{
    return flush_common(c, n_streams, mp3, mp3_stride, lengths, true, (cudaStream_t)stream);
}

extern "C" int mp3gpu_flush_mp3_dev(mp3gpu_ctx *c, int n_streams, uint8_t *mp3, long mp3_stride, long *lengths, void *stream)
{
    return flush_common(c, n_streams, mp3, mp3_stride, lengths, false, (cudaStream_t)stream);
}

extern "C" int mp3gpu_encode_frames(mp3gpu_ctx *c, const int16_t *pcm, int n_streams, int n_frames, int16_t *ix, mp3gpu_gr_info *gi,
                                    uint8_t *sf, mp3gpu_frame_out *fo, void *stream)
{
    if (!c) return fail(MP3GPU_EINVAL, "null ctx");
    DEV_GUARD(c);
    return encode_common(c, pcm, n_streams, n_frames, ix, gi, sf, fo, stream, true);
}

extern "C" int mp3gpu_encode_frames_dev(mp3gpu_ctx *c, const int16_t *pcm, int n_streams, int n_frames, int16_t *ix, mp3gpu_gr_info *gi,
                                        uint8_t *sf, mp3gpu_frame_out *fo, void *stream)
{
    if (!c) return fail(MP3GPU_EINVAL, "null ctx");
    DEV_GUARD(c);
    return encode_common(c, pcm, n_streams, n_frames, ix, gi, sf, fo, stream, false);
}

extern "C" int mp3gpu_filter_subband_batch(mp3gpu_ctx *c, const int16_t *pcm, int n_streams, int n_frames, double *sb, void *stream)
{
    int rc = check_shape(c, n_streams, n_frames);
    if (rc) return rc;
    if (!pcm || !sb) return fail(MP3GPU_EINVAL, "null pointer");
    DEV_GUARD(c);
    cudaStream_t q = (cudaStream_t)stream;
    if ((rc = stage_pcm(c, c->pcm_fb, pcm, n_streams, n_frames, cudaMemcpyDeviceToDevice, q))) return rc;
    if ((rc = launch_front(c, c->pcm_fb.buf, nullptr, n_streams, n_frames, nullptr, sb, false, q))) return rc;
    return roll_pcm(c, c->pcm_fb, n_streams, n_frames, q);
}

extern "C" int mp3gpu_subband_mdct_batch(mp3gpu_ctx *c, const int16_t *pcm, const mp3gpu_psy_out *psy, int n_streams, int n_frames,
                                         double *xr, void *stream)
{
    int rc = check_shape(c, n_streams, n_frames);
    if (rc) return rc;
    if (!pcm || !psy || !xr) return fail(MP3GPU_EINVAL, "null pointer");
    DEV_GUARD(c);
    cudaStream_t q = (cudaStream_t)stream;
    if ((rc = stage_pcm(c, c->pcm_fb, pcm, n_streams, n_frames, cudaMemcpyDeviceToDevice, q))) return rc;
    if ((rc = launch_front(c, c->pcm_fb.buf, (const PsyOut *)psy, n_streams, n_frames, xr, nullptr, true, q))) return rc;
    return roll_pcm(c, c->pcm_fb, n_streams, n_frames, q);
}

extern "C" int mp3gpu_mdct_sub_batch(mp3gpu_ctx *c, const double *sb, const mp3gpu_psy_out *psy, int n_streams, int n_frames, double *xr,
                                     void *stream)
{
    int rc = check_shape(c, n_streams, n_frames);
    if (rc) return rc;
    if (!sb || !psy || !xr) return fail(MP3GPU_EINVAL, "null pointer");
    DEV_GUARD(c);
    cudaStream_t q = (cudaStream_t)stream;
    if (!c->d_sb_prev) {
        const size_t n = (size_t)c->cfg.max_streams * c->cfg.n_ch * 576;
        if ((rc = dalloc(&c->d_sb_prev, n))) return rc;
        CU(cudaMemsetAsync(c->d_sb_prev, 0, n * sizeof(double), q));
    }
    const long chans = (long)n_streams * c->cfg.n_ch;
    k_mdct<<<(unsigned)((chans + FRONT_WARPS - 1) / FRONT_WARPS), FRONT_WARPS * 32, FRONT_WARPS * sizeof(FrontWarpSmem), q>>>(
        sb, (const PsyOut *)psy, c->d_sb_prev, n_streams, c->cfg.n_ch, 2 * n_frames, xr);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}

extern "C" int mp3gpu_L3psycho_anal_batch(mp3gpu_ctx *c, const int16_t *pcm, int n_streams, int n_frames, mp3gpu_psy_out *psy, void *stream)
{
    int rc = check_shape(c, n_streams, n_frames);
    if (rc) return rc;
    if (!pcm || !psy) return fail(MP3GPU_EINVAL, "null pointer");
    DEV_GUARD(c);
    cudaStream_t q = (cudaStream_t)stream;
    if ((rc = stage_pcm(c, c->pcm_psy, pcm, n_streams, n_frames, cudaMemcpyDeviceToDevice, q))) return rc;
    if ((rc = launch_psy(c, c->pcm_psy.buf, n_streams, n_frames, (PsyOut *)psy, q))) return rc;
    return roll_pcm(c, c->pcm_psy, n_streams, n_frames, q);
}

extern "C" int mp3gpu_iteration_loop_batch(mp3gpu_ctx *c, const double *xr, const mp3gpu_psy_out *psy, int n_streams, int n_frames,
                                           int16_t *ix, mp3gpu_gr_info *gi, uint8_t *sf, mp3gpu_frame_out *fo, void *stream)
{
    int rc = check_shape(c, n_streams, n_frames);
    if (rc) return rc;
    if (!xr || !psy || !ix || !gi || !sf || !fo) return fail(MP3GPU_EINVAL, "null pointer");
    DEV_GUARD(c);
    return launch_rate_loop(c, xr, (const PsyOut *)psy, n_streams, n_frames, ix, (GrInfoOut *)gi, sf, (FrameOut *)fo, (cudaStream_t)stream);
}

extern "C" int mp3gpu_quantize_count_batch(mp3gpu_ctx *c, const double *xr_abs, const int *q, const int *block_type, int n, int16_t *ix,
                                           mp3gpu_gr_info *gi, int *bits, void *stream)
{
    if (!c) return fail(MP3GPU_EINVAL, "null ctx");
    if (n < 1 || !xr_abs || !q || !block_type || !ix || !gi || !bits) return fail(MP3GPU_EINVAL, "bad argument");
    DEV_GUARD(c);
    k_quantize_count<<<(unsigned)((n + RL_WARPS - 1) / RL_WARPS), RL_WARPS * 32, RL_SMEM_BYTES, (cudaStream_t)stream>>>(
        c->d_rate_tab, xr_abs, q, block_type, n, ix, (GrInfoOut *)gi, bits, 0);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}

extern "C" int mp3gpu_count_bits_batch(mp3gpu_ctx *c, const int16_t *ix, const int *block_type, int n, mp3gpu_gr_info *gi, int *bits,
                                       void *stream)
{
    if (!c) return fail(MP3GPU_EINVAL, "null ctx");
    if (n < 1 || !ix || !block_type || !gi || !bits) return fail(MP3GPU_EINVAL, "bad argument");
    DEV_GUARD(c);
    k_quantize_count<<<(unsigned)((n + RL_WARPS - 1) / RL_WARPS), RL_WARPS * 32, RL_SMEM_BYTES, (cudaStream_t)stream>>>(
        c->d_rate_tab, nullptr, nullptr, block_type, n, const_cast<int16_t *>(ix), (GrInfoOut *)gi, bits, 1);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}

#include "legacy_shim.cuh"
