// front_tile.cuh — production kernel for the fused polyphase filterbank + MDCT + alias reduction
// (replaces window_subband()/filter_subband(), /root/reference/src/encode.c:287-409, and
// mdct_sub()/mdct(), mdct.c:25-198, for mp3gpu_encode_frames / mp3gpu_subband_mdct_batch).
//
// One CTA of 288 threads encodes a TILE of up to 15 consecutive granules of one (stream, channel),
// plus one warm-up granule in front (the MDCT needs the previous granule's subband samples).
// 16 granules = 288 polyphase slots = 9 warps x 32 slots, so every stage runs with full warps:
//
//   stage 0  PCM tile (int16) HBM -> shared memory, 16-byte coalesced loads
//   stage A  windowing + 8-fold (encode.c:310-311,392-398).  lane = tap index i: each thread runs two
//            8-tap FIRs (i, i+32) with the window taps in registers and a register sliding window of
//            samples (1 shared load per output instead of 8); the fold to ysum/ysub is two shuffles.
//   stage B  32x31 matrixing (encode.c:399-408).  thread = SLOT: the 31 folded inputs live in
//            registers, the cosine matrix is read as immediate constant-bank operands of fully
//            unrolled DMUL/DADD (no loads in the inner loop), results overwrite the thread's own row.
//   stage C  MDCT (mdct.c:171-198), three warps per granule (6 outputs each), lane = band, cosine tables as immediates.
//   stage D  alias butterflies (mdct.c:83-91) in shared memory, then coalesced 128-bit stores of xr.
//
// Arithmetic: every sum is evaluated in the reference's order with unfused IEEE mul/add, so subband
// samples are bit-identical to the reference and xr is bit-identical to the oracle (the reference's
// hand-unrolled type-0 MDCT differs only in summation order, <= 2e-14 relative).
// The bound of this kernel is the FP64 pipe (about 100 k FP64 instructions per granule-channel against
// 5764 algorithmic bytes), see DESIGN.md.
#pragma once
#include <cuda_runtime.h>

#include "tables.h"

namespace mp3gpu {

#define FT_G 15                         // output granules per tile
#define FT_SLOTS (18 * (FT_G + 1))      // 288
#define FT_THREADS FT_SLOTS
#define FT_WARPS (FT_THREADS / 32)      // 9
#define FT_ROW 33                       // doubles per slot row (odd: conflict-free for lane = slot)
#define FT_PCM (480 + 32 * FT_SLOTS)    // 9696 samples

struct FrontTileSmem {
    double rows[FT_SLOTS * FT_ROW];     // stage A: ysum/ysub/y16; stage B: subband samples (in place);
                                        // stage D: xr staging in the rows of the previous granule
    double window[512];                 // Table C.1                        } loaded once per CTA: the CTA is persistent and
    double tab[32 * 32 + 18 * 36];      // am[32][32], cos_l[18][36]        } walks over tiles
    short pcm[FT_PCM];
    int bt[FT_G + 1];                   // block types of the tile's granules (L3psycho_anal's decision)
};
// layout of tab[] (doubles)
#define FT_TAB_AM 0                     // am[32][32]
#define FT_TAB_COS (32 * 32)            // cos_l[18][36]
#define FT_TAB_END (FT_TAB_COS + 18 * 36)
#ifndef FT_TABLES_IN_SMEM
#define FT_TABLES_IN_SMEM 1
#endif

__constant__ FrontTables c_front;  // defined here: this header is included by exactly one translation unit (mp3gpu.cu)
// the same tables in global memory: per-thread-indexed reads of __constant__ data serialise in the constant cache, so the
// cooperative copies into shared memory read this copy with coalesced 16-byte loads instead
__device__ __align__(16) FrontTables g_front;

// ---- stage B: s[sb] = y16 + sum_j am[sb][j] * ys[j], j ascending (encode.c:399-408) ----------------
// Eight rows (independent accumulation chains) per call; the row group is a RUN-TIME offset so that the 496 FP64
// instructions exist once and loop four times: fully unrolled over all 32 rows the stage streamed 38 KB of code per pass
// and stalled on instruction fetch (ncu: stall_no_instruction 1.4 warps per issue).
__device__ __forceinline__ void ft_matrix_rows8(const double (&ys)[32], double *row, int sb0, bool odd_slot, const double *tab)
{
    double s[8];
#if FT_TABLES_IN_SMEM
    const double2 *a = reinterpret_cast<const double2 *>(tab + FT_TAB_AM + sb0 * 32);   // uniform address: broadcast loads
#pragma unroll
    for (int k = 0; k < 8; k++) s[k] = ys[31];
#pragma unroll
    for (int j = 0; j < 30; j += 2)
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const double2 c = a[16 * k + (j >> 1)];
            s[k] = __dadd_rn(s[k], __dmul_rn(c.x, ys[j]));
            s[k] = __dadd_rn(s[k], __dmul_rn(c.y, ys[j + 1]));
        }
#pragma unroll
    for (int k = 0; k < 8; k++) s[k] = __dadd_rn(s[k], __dmul_rn(a[16 * k + 15].x, ys[30]));
#else
#pragma unroll
    for (int k = 0; k < 8; k++) {
        s[k] = ys[31];
#pragma unroll
        for (int j = 0; j < 31; j++) s[k] = __dadd_rn(s[k], __dmul_rn(c_front.am[sb0 + k][j], ys[j]));
    }
#endif
#pragma unroll
    for (int k = 0; k < 8; k++) {
        // mdct.c:57-60: odd band, odd time slot -> * -1 (applied here, the raw value is never needed); sb0 is a multiple of 8
        const double v = ((k & 1) && odd_slot) ? __dmul_rn(s[k], -1.0) : s[k];
        row[sb0 + k] = v;
    }
}

// ---- stage C: MDCT of one band held in registers, one THIRD of the outputs per call ------------------
// A granule's 18 outputs per band are computed by three warps (6 outputs each): 15 granules x 3 = 45 work items
// = 5 full rounds of the 9 warps, and 6 independent accumulation chains per thread.  The window products
// fin[k] = win[bt][k] * in[k] (mdct.c:190-192) are recomputed by each of the three warps (36 of 468 operations).
template <int WIN>
__device__ __forceinline__ void ft_window(double (&fin)[36])
{
    // the window is a template parameter so that every table read has a compile-time address (a register-indexed
    // __constant__ read goes through the ADU)
#pragma unroll
    for (int k = 0; k < 36; k++) fin[k] = __dmul_rn(c_front.win[WIN][k], fin[k]);
}

__device__ __forceinline__ void ft_mdct_long6(const double (&fin)[36], double (&out)[6], int m0, const double *tab)
{
#pragma unroll
    for (int m = 0; m < 6; m++) out[m] = 0.0;
#if FT_TABLES_IN_SMEM
    const double *ct = tab + FT_TAB_COS + m0 * 36;           // one copy of the code for the three thirds (run-time m0)
#pragma unroll
    for (int k = 0; k < 36; k += 2)                                                 // mdct.c:193-198, k ascending per output
#pragma unroll
        for (int m = 0; m < 6; m++) {
            const double2 c = *reinterpret_cast<const double2 *>(ct + m * 36 + k);
            out[m] = __dadd_rn(out[m], __dmul_rn(fin[k], c.x));
            out[m] = __dadd_rn(out[m], __dmul_rn(fin[k + 1], c.y));
        }
#else
#pragma unroll
    for (int k = 0; k < 36; k++)
#pragma unroll
        for (int m = 0; m < 6; m++) out[m] = __dadd_rn(out[m], __dmul_rn(fin[k], c_front.cos_l[m0 + m][k]));
#endif
}

template <int L>
__device__ __forceinline__ void ft_mdct_short6(const double (&in)[36], double (&out)[6])
{
#pragma unroll
    for (int m = 0; m < 6; m++) out[m] = 0.0;
#pragma unroll
    for (int k = 0; k < 12; k++) {                                                  // mdct.c:171-185, window L
        const double f = __dmul_rn(c_front.win[2][k], in[k + 6 * L + 6]);
#pragma unroll
        for (int m = 0; m < 6; m++) out[m] = __dadd_rn(out[m], __dmul_rn(f, c_front.cos_s[m][k]));
    }
}

// ---- bulk asynchronous copy (TMA engine, SASS UBLKCP) of the contiguous PCM tile into shared memory -----------------------
// One elected thread posts the copy and the byte count on an mbarrier; every thread waits on the barrier's phase.  The
// tile is one contiguous, 16-byte aligned run of a PCM row, so the 1-D bulk form is enough (no tensor map).
#ifndef FT_TMA
#define FT_TMA 1      // A/B (profiles/r02_front_variants.md): 1 = cp.async.bulk, 0 = 16-byte loads by all threads
#endif
__device__ __forceinline__ unsigned ft_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ft_mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ft_smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void ft_bulk_load(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ft_smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(ft_smem_u32(dst)), "l"(src), "r"(bytes), "r"(ft_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ft_mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra WAIT_%=;\n\t}"
                 ::"r"(ft_smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void ft_group_barrier(int group)   // the three warps of one granule
{
    asm volatile("bar.sync %0, 96;" ::"r"(group + 1) : "memory");
}

// pcm_rows: [n_streams*n_ch] rows of `row` samples, HIST samples of history first (see mp3gpu.cu)
// One CTA per tile (with FT_PERSISTENT: 2 CTAs per SM, CTA b takes tiles b, b + gridDim.x, ... of the n_tiles_total =
// streams x channels x tiles-per-row tiles).  The window and cosine tables have their own shared-memory area and are
// fetched together with the PCM tile.
__global__ void __launch_bounds__(FT_THREADS, 2)
k_front_tile(const short *__restrict__ pcm_rows, long stream_stride, long ch_stride, int hist, int n_streams, int n_ch, int n_gran,
             long n_tiles_total, const int *__restrict__ nfr, const PsyOut *__restrict__ psy, double *__restrict__ xr)
{
    extern __shared__ __align__(16) unsigned char ft_smem_raw[];
    FrontTileSmem &M = *reinterpret_cast<FrontTileSmem *>(ft_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_tiles = (n_gran + FT_G - 1) / FT_G;
    const double *tab = M.tab;
#if FT_TMA && !FT_PERSISTENT
    // the PCM tile is on its way (TMA engine) while the threads fetch the tables
    __shared__ unsigned long long ft_bar;
    {
        const long bid0 = blockIdx.x;
        const int t0 = (int)(bid0 % n_tiles), ch0 = (int)((bid0 / n_tiles) % n_ch);
        const long s0 = bid0 / ((long)n_tiles * n_ch);
        const int live0 = nfr ? min(n_gran, 2 * nfr[s0]) : n_gran, ng0 = min(FT_G, live0 - t0 * FT_G);
        if (ng0 <= 0) return;                                       // CTA-uniform: the stream ended before this tile
        if (tid == 0) {
            ft_mbar_init(&ft_bar, 1);
            const short *src0 = pcm_rows + s0 * stream_stride + ch0 * ch_stride + hist + 576L * (t0 * FT_G - 1) - 480;
            ft_bulk_load(M.pcm, src0, (unsigned)((480 + 32 * 18 * (ng0 + 1)) * sizeof(short)), &ft_bar);
        }
    }
#endif
    {
        if (tid < 256) reinterpret_cast<double2 *>(M.window)[tid] = reinterpret_cast<const double2 *>(g_front.window)[tid];
        double2 *t2 = reinterpret_cast<double2 *>(M.tab);
        const double2 *am2 = reinterpret_cast<const double2 *>(&g_front.am[0][0]), *cos2 = reinterpret_cast<const double2 *>(&g_front.cos_l[0][0]);
        for (int i = tid; i < 32 * 32 / 2; i += FT_THREADS) t2[FT_TAB_AM / 2 + i] = am2[i];
        for (int i = tid; i < 18 * 36 / 2; i += FT_THREADS) t2[FT_TAB_COS / 2 + i] = cos2[i];
    }
#ifndef FT_PERSISTENT
#define FT_PERSISTENT 0     // A/B: the tile loop costs registers (spills under the 96-register cap): 44 ms vs 28 ms per step
#endif
#if FT_PERSISTENT
  for (long bid = blockIdx.x; bid < n_tiles_total; bid += gridDim.x) {
#else
  {
    const long bid = blockIdx.x;
#endif
    const int t = (int)(bid % n_tiles);
    const int ch = (int)((bid / n_tiles) % n_ch);
    const long s = bid / ((long)n_tiles * n_ch);
    const int g_first = t * FT_G;
    const int n_live = nfr ? min(n_gran, 2 * nfr[s]) : n_gran;      // granules of this stream in this call (mp3gpu_set_stream_frames)
    const int ng = min(FT_G, n_live - g_first);
#if FT_PERSISTENT
    if (ng <= 0) continue;
#else
    if (ng <= 0) return;                                            // CTA-uniform: the stream ended before this tile
#endif
    const int n_slots = 18 * (ng + 1);
    const int n_chunks = (n_slots + 31) >> 5;

    // ---- stage 0: PCM tile.  smem sample j <-> stream time 576*(g_first-1) - 480 + j ---------------
    {
        const int n_valid = 480 + 32 * n_slots;           // multiple of 8
        uint4 *dst4 = reinterpret_cast<uint4 *>(M.pcm);
#if FT_TMA && !FT_PERSISTENT
        for (int i = n_valid / 8 + tid; i < FT_PCM / 8; i += FT_THREADS) dst4[i] = make_uint4(0, 0, 0, 0);   // a short last tile
#else
        const short *src = pcm_rows + s * stream_stride + ch * ch_stride + hist + 576L * (g_first - 1) - 480;
        const uint4 *src4 = reinterpret_cast<const uint4 *>(src);
        for (int i = tid; i < FT_PCM / 8; i += FT_THREADS)
            dst4[i] = (i < n_valid / 8) ? src4[i] : make_uint4(0, 0, 0, 0);
#endif
        if (tid < ng) M.bt[tid] = psy[((s * n_gran + g_first + tid) * (long)n_ch + ch)].block_type;
    }
#if FT_TMA && !FT_PERSISTENT
    __syncthreads();                                      // thread 0's mbarrier.init is visible to every waiter
    ft_mbar_wait(&ft_bar, 0);
#endif
    __syncthreads();

    // ---- stage A: lane = tap i (and i+32); slots of one parity form a sliding 8-tap FIR -------------
    for (int c = warp; c < n_chunks; c += FT_WARPS) {
        double w0[8], w1[8];
#pragma unroll
        // encode.c:306 divides the sample by SCALE = 32768 before the window multiply (encode.c:310); a power of two moves through
        // the rounding of the product unchanged, so the taps carry it: (x / 32768) * w == x * (w / 32768) bit for bit (no
        // product comes near the denormal range: |w| >= 4.77e-7 where it is not zero)
        for (int j = 0; j < 8; j++) { w0[j] = __dmul_rn(M.window[lane + 64 * j], 1.0 / 32768); w1[j] = __dmul_rn(M.window[lane + 32 + 64 * j], 1.0 / 32768); }
        const int src_lane = (32 - lane) & 31;
#pragma unroll 1
        for (int p = 0; p < 2; p++) {
            // slot n = 32c + 2k + p; newest sample of tap i: pcm[480 + 32n + 31 - i]
            const int base0 = 480 + 32 * (32 * c + p) + 31 - lane;      // k = 0, i = lane
            double h0[8], h1[8];                                         // h[j] = sample for tap j of the CURRENT slot
#pragma unroll
            for (int j = 1; j < 8; j++) {
                h0[j] = (double)M.pcm[base0 - 64 * j];
                h1[j] = (double)M.pcm[base0 - 32 - 64 * j];
            }
#pragma unroll
            for (int k = 0; k < 16; k++) {
                h0[0] = (double)M.pcm[base0 + 64 * k];
                h1[0] = (double)M.pcm[base0 - 32 + 64 * k];
                double y0 = __dmul_rn(h0[0], w0[0]), y1 = __dmul_rn(h1[0], w1[0]);   // encode.c:310-311,392-396
#pragma unroll
                for (int j = 1; j < 8; j++) {
                    y0 = __dadd_rn(y0, __dmul_rn(h0[j], w0[j]));
                    y1 = __dadd_rn(y1, __dmul_rn(h1[j], w1[j]));
                }
#pragma unroll
                for (int j = 7; j > 0; j--) { h0[j] = h0[j - 1]; h1[j] = h1[j - 1]; }
                const double a0 = __shfl_sync(0xffffffffu, y0, src_lane);            // y[32 - i]
                const double a1 = __shfl_sync(0xffffffffu, y1, src_lane);            // y[64 - i]
                double *row = M.rows + (size_t)(32 * c + 2 * k + p) * FT_ROW;
                if (lane == 0) row[0] = __dadd_rn(y0, y1);                            // ysum[0] = y[0] + y[32]
                else if (lane < 16) {
                    row[lane] = __dadd_rn(y0, a0);                                    // ysum[i] = y[i] + y[32-i]
                    row[15 + lane] = __dsub_rn(y1, a1);                               // ysub[i-1] = y[32+i] - y[64-i]
                } else if (lane == 16) row[31] = y0;                                  // y[16]
            }
        }
    }
    // chunk c = warp of stage A is exactly the 32 slots the same warp transforms in stage B (n_chunks <= FT_WARPS): a warp
    // barrier is enough, and warps drift from the FIR stage into the matrixing stage instead of meeting at a CTA barrier
    __syncwarp();
    // ---- stage B: thread = slot ---------------------------------------------------------------------
    if (tid < n_chunks * 32) {
        double *row = M.rows + (size_t)tid * FT_ROW;
        double ys[32];
#pragma unroll
        for (int j = 0; j < 32; j++) ys[j] = row[j];
        const bool odd_slot = ((tid % 18) & 1) != 0;
#pragma unroll 1
        for (int sb0 = 0; sb0 < 32; sb0 += 8) ft_matrix_rows8(ys, row, sb0, odd_slot, tab);
    }
    __syncthreads();

    // ---- stage C + D: (granule, third) per warp, lane = band ------------------------------------------
    // Warps 3j, 3j+1, 3j+2 share a granule.  Outputs are staged in the rows of the PREVIOUS granule, which are
    // inputs of this round only: every warp loads its inputs, one CTA barrier, then compute / butterflies / store
    // with barriers among the three warps of the granule.
    const long gc_stride = n_ch;
    const int third = warp % 3, group = warp / 3;
    for (int r = 0; 3 * r < ng; r++) {
        const int gl = 1 + 3 * r + group;                  // local granule index, 1..ng (0 is the warm-up granule)
        const bool active = gl <= ng;
        double in[36], out[6];
        int bt = 0;
        if (active) {
            const double *p = M.rows + (size_t)(18 * (gl - 1)) * FT_ROW + lane;
#pragma unroll
            for (int k = 0; k < 36; k++) in[k] = p[k * FT_ROW];
            bt = M.bt[gl - 1];
        }
        __syncthreads();                                    // every warp has its inputs: previous granules' rows are free
        if (active) {
            double *st = M.rows + (size_t)(18 * (gl - 1)) * FT_ROW;   // 576 doubles in the previous granule's rows
            if (bt == 2) {
                if (third == 0) ft_mdct_short6<0>(in, out); else if (third == 1) ft_mdct_short6<1>(in, out); else ft_mdct_short6<2>(in, out);
#pragma unroll
                for (int m = 0; m < 6; m++) st[lane * 18 + 3 * m + third] = out[m];
            } else {
                if (bt == 0) ft_window<0>(in); else if (bt == 1) ft_window<1>(in); else ft_window<3>(in);
                ft_mdct_long6(in, out, 6 * third, tab);
#pragma unroll
                for (int m = 0; m < 6; m++) st[lane * 18 + 6 * third + m] = out[m];
            }
            ft_group_barrier(group);
            if (bt != 2 && lane < 31) {                                                 // mdct.c:83-91, butterflies k = third, third+3, third+6
                for (int k = third; k < 8; k += 3) {
                    const double a = st[lane * 18 + 17 - k], b = st[(lane + 1) * 18 + k];
                    const double bu = __dadd_rn(__dmul_rn(a, c_front.cs[k]), __dmul_rn(b, c_front.ca[k]));
                    const double bd = __dsub_rn(__dmul_rn(b, c_front.cs[k]), __dmul_rn(a, c_front.ca[k]));
                    st[lane * 18 + 17 - k] = bu;
                    st[(lane + 1) * 18 + k] = bd;
                }
            }
            ft_group_barrier(group);
            double2 *dst = reinterpret_cast<double2 *>(xr + ((s * n_gran + g_first + gl - 1) * gc_stride + ch) * 576);
            const double2 *src = reinterpret_cast<const double2 *>(st);
#pragma unroll
            for (int i = 0; i < 3; i++) dst[lane + 32 * (3 * third + i)] = src[lane + 32 * (3 * third + i)];
        }
    }
    __syncthreads();        // the next tile's stages overwrite pcm[] and rows[]
  }
}

}  // namespace mp3gpu
