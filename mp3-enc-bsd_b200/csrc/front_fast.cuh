// front_fast.cuh — tolerance-path variants of the fused polyphase filterbank + MDCT + alias reduction kernel
// (window_subband()/filter_subband(), /root/reference/src/encode.c:287-409; mdct_sub()/mdct(), mdct.c:25-198).
//
// front_tile.cuh evaluates every sum in the reference's order with unfused IEEE mul/add: bit-identical, and FP64-pipe bound
// at ~100 k FP64 instructions per granule-channel.  BASELINE's north star allows <= 1e-12 (FP64) and <= 1e-5 (FP32) relative
// error for these stages; the variants here spend that tolerance on arithmetic:
//
//   MP3GPU_FRONT_FMA  (T = double)  fused multiply-add everywhere; the 32 x 31 matrixing stays direct-form on the reference's
//                     coefficients (create_ana_filter rounds them to 9 decimals, encode.c:331-345: a fast DCT on exact
//                     cosines would differ from the reference by ~1e-9); the 36 -> 18 MDCT is folded into an 18-point
//                     DCT-IV (time-domain aliasing: u[j] = -f[26-j] - f[27+j] for j < 9, f[j-9] - f[26-j] for j >= 9;
//                     X[m] = sum_j u[j] cos(pi/72 (2j+1)(2m+1)) / 9): 378 instead of 1332 FP64 operations per band.
//                     ~40 k FP64 instructions per granule-channel; error ~1e-15 relative to the granule's largest value.
//   MP3GPU_FRONT_FP32 (T = float)   the same dataflow in FP32 (error ~1e-6), the matrixing as Lee's fast 32-point DCT-III
//                     (209 fused operations per slot instead of 992; its exact cosines are 5e-10 from the reference's
//                     rounded ones, far inside the FP32 tolerance), xr stored as float (3460 algorithmic bytes per
//                     granule-channel instead of 5764) for a rate loop that reads float spectra.
//
// Same tiling as front_tile.cuh: one CTA of 288 threads per tile of 15 granules of one (stream, channel) + 1 warm-up granule.
#pragma once
#include <cuda_runtime.h>

#include "front_tile.cuh"

namespace mp3gpu {

#define FF_C4_ROW_ 20
// FP32 copies of the small tables the kernel reads with compile-time indices (a double constant read would pay an F2F each)
struct FrontTablesF {
    float win[4][36];
    float dct4_s[6][6];
    float ca[8], cs[8];
    float window[512];
    float dct4_l[18][FF_C4_ROW_];
};
__constant__ FrontTablesF c_front_f;
__device__ __align__(16) FrontTablesF g_front_f;

template <class T> struct FastTab;
template <> struct FastTab<double> {
    static __device__ __forceinline__ double win(int w, int k) { return c_front.win[w][k]; }
    static __device__ __forceinline__ double dct4_s(int m, int j) { return c_front.dct4_s[m][j]; }
    static __device__ __forceinline__ double ca(int k) { return c_front.ca[k]; }
    static __device__ __forceinline__ double cs(int k) { return c_front.cs[k]; }
    static __device__ __forceinline__ double window(int i) { return g_front.window[i]; }
    static __device__ __forceinline__ double dct4_l(int m, int j) { return g_front.dct4_l[m][j]; }
    static __device__ __forceinline__ double c4c(int m, int j) { return c_front.dct4_l[m][j]; }     // constant bank
};
template <> struct FastTab<float> {
    static __device__ __forceinline__ float win(int w, int k) { return c_front_f.win[w][k]; }
    static __device__ __forceinline__ float dct4_s(int m, int j) { return c_front_f.dct4_s[m][j]; }
    static __device__ __forceinline__ float ca(int k) { return c_front_f.ca[k]; }
    static __device__ __forceinline__ float cs(int k) { return c_front_f.cs[k]; }
    static __device__ __forceinline__ float window(int i) { return g_front_f.window[i]; }
    static __device__ __forceinline__ float dct4_l(int m, int j) { return g_front_f.dct4_l[m][j]; }
    static __device__ __forceinline__ float c4c(int m, int j) { return c_front_f.dct4_l[m][j]; }
};

// Coefficients as constant-bank operands instead of shared-memory broadcast loads (ncu, profiles/r02_front_variants.md: the
// shared-memory pipe, not the FP64 pipe, bounds these kernels — 56 % of the wavefronts of the FMA variant were coefficient
// loads).  A warp-uniform run-time index compiles to LDCU (uniform constant load) + FMA with a uniform-register operand; a
// compile-time index to a c[bank][imm] operand.
#ifndef FF_COEF_CONST
#define FF_COEF_CONST 1
#endif

template <class T> struct FastArith;
template <> struct FastArith<double> {
    static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
};
template <> struct FastArith<float> {
    static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
};

#define FF_C4_ROW FF_C4_ROW_              // padded row of the 18-point DCT-IV matrix (16-byte aligned rows for float4 / double2 loads)

template <class T, bool LEE>
struct FrontFastSmem {
    T rows[FT_SLOTS * FT_ROW];
    T window[512];
    T am[LEE ? 4 : 32 * FT_ROW];         // direct-form matrixing coefficients (unused with Lee's DCT; rows padded for the DMMA variant)
    T c4[18 * FF_C4_ROW];                // cos(pi/72 (2j+1)(2m+1)) / 9
    short pcm[FT_PCM];
    int bt[FT_G + 1];
};

// ---- Lee's fast DCT-III, in registers: X[k] = sum_n x[n] cos((2k+1) n pi / 2N) -----------------------------------------
//   g[n] = x[2n], h[n] = x[2n+1] + x[2n-1] (x[-1] = 0);  X[k] = G[k] + H[k] / (2 cos((2k+1) pi / 2N)),  X[N-1-k] = G[k] - ...
// The recursion is unrolled at compile time; the 31 reciprocal-cosine constants are immediates.
template <int N>
__device__ __forceinline__ constexpr float lee_coef(int k)         // 1 / (2 cos((2k+1) pi / 2N))
{
    constexpr float c2[1] = {0.70710678118654746f};
    constexpr float c4[2] = {0.54119610014619701f, 1.3065629648763764f};
    constexpr float c8[4] = {0.50979557910415918f, 0.60134488693504529f, 0.89997622313641557f, 2.5629154477415055f};
    constexpr float c16[8] = {0.50241928618815568f, 0.52249861493968885f, 0.56694403481635769f, 0.64682178335999008f,
                              0.7881546234512502f, 1.0606776859903471f, 1.7224470982383342f, 5.1011486186891553f};
    constexpr float c32[16] = {0.50060299823519627f, 0.50547095989754365f, 0.51544730992262455f, 0.53104259108978413f,
                               0.55310389603444454f, 0.58293496820613389f, 0.62250412303566482f, 0.67480834145500568f,
                               0.74453627100229858f, 0.83934964541552681f, 0.97256823786196078f, 1.1694399334328847f,
                               1.4841646163141662f, 2.0577810099534108f, 3.407608418468719f, 10.190008123548033f};
    return N == 2 ? c2[k] : N == 4 ? c4[k & 1] : N == 8 ? c8[k & 3] : N == 16 ? c16[k & 7] : c32[k & 15];
}

template <int N, class T>
__device__ __forceinline__ void lee_dct3(T (&x)[N])
{
    if constexpr (N == 1) {
        return;
    } else {
        T g[N / 2], h[N / 2];
#pragma unroll
        for (int n = 0; n < N / 2; n++) {
            g[n] = x[2 * n];
            h[n] = (n == 0) ? x[1] : FastArith<T>::add(x[2 * n + 1], x[2 * n - 1]);
        }
        lee_dct3<N / 2, T>(g);
        lee_dct3<N / 2, T>(h);
#pragma unroll
        for (int k = 0; k < N / 2; k++) {
            const T c = (T)lee_coef<N>(k);
            x[k] = FastArith<T>::fma(h[k], c, g[k]);
            x[N - 1 - k] = FastArith<T>::fma(h[k], -c, g[k]);
        }
    }
}

// ---- stage B, direct form with FMA on the reference's (rounded) coefficients: eight rows per call ------------------------
template <class T>
__device__ __forceinline__ void ff_matrix_rows8(const T (&ys)[32], T *row, int sb0, bool odd_slot, const T *am)
{
    T s[8];
#pragma unroll
    for (int k = 0; k < 8; k++) s[k] = ys[31];
#if FF_COEF_CONST
    if constexpr (sizeof(T) == 8) {
        (void)am;
#pragma unroll
        for (int j = 0; j < 31; j++)
#pragma unroll
            for (int k = 0; k < 8; k++) s[k] = FastArith<T>::fma(c_front.am[sb0 + k][j], ys[j], s[k]);     // sb0: loop counter, warp-uniform
    } else
#endif
    if constexpr (sizeof(T) == 8) {
        const double2 *a = reinterpret_cast<const double2 *>(am + sb0 * 32);
#pragma unroll
        for (int j = 0; j < 30; j += 2)
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const double2 c = a[16 * k + (j >> 1)];
                s[k] = FastArith<T>::fma(c.x, ys[j], s[k]);
                s[k] = FastArith<T>::fma(c.y, ys[j + 1], s[k]);
            }
#pragma unroll
        for (int k = 0; k < 8; k++) s[k] = FastArith<T>::fma(a[16 * k + 15].x, ys[30], s[k]);
    } else {
        const float4 *a = reinterpret_cast<const float4 *>(am + sb0 * 32);
#pragma unroll
        for (int j = 0; j < 28; j += 4)
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float4 c = a[8 * k + (j >> 2)];
                s[k] = FastArith<T>::fma(c.x, ys[j], s[k]);
                s[k] = FastArith<T>::fma(c.y, ys[j + 1], s[k]);
                s[k] = FastArith<T>::fma(c.z, ys[j + 2], s[k]);
                s[k] = FastArith<T>::fma(c.w, ys[j + 3], s[k]);
            }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float4 c = a[8 * k + 7];
            s[k] = FastArith<T>::fma(c.x, ys[28], s[k]);
            s[k] = FastArith<T>::fma(c.y, ys[29], s[k]);
            s[k] = FastArith<T>::fma(c.z, ys[30], s[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) row[sb0 + k] = ((k & 1) && odd_slot) ? -s[k] : s[k];      // mdct.c:57-60
}

// ---- stage C: windowing + time-domain aliasing fold of one band, then six outputs of the 18-point DCT-IV ------------------
template <int WIN, class T>
__device__ __forceinline__ void ff_fold_long(const T (&in)[36], T (&u)[18])
{
    T f[36];
#pragma unroll
    for (int k = 0; k < 36; k++) f[k] = FastArith<T>::mul(FastTab<T>::win(WIN, k), in[k]);
#pragma unroll
    for (int j = 0; j < 9; j++) u[j] = FastArith<T>::sub(-f[26 - j], f[27 + j]);
#pragma unroll
    for (int j = 9; j < 18; j++) u[j] = FastArith<T>::sub(f[j - 9], f[26 - j]);
}

template <int M0, class T>
__device__ __forceinline__ void ff_dct4_6c(const T (&u)[18], T (&out)[6])          // coefficients: c[bank][imm] operands
{
#pragma unroll
    for (int m = 0; m < 6; m++) out[m] = (T)0;
#pragma unroll
    for (int j = 0; j < 18; j++)
#pragma unroll
        for (int m = 0; m < 6; m++) out[m] = FastArith<T>::fma(u[j], FastTab<T>::c4c(M0 + m, j), out[m]);
}

template <class T>
__device__ __forceinline__ void ff_dct4_6(const T (&u)[18], T (&out)[6], int m0, const T *c4)
{
#pragma unroll
    for (int m = 0; m < 6; m++) out[m] = (T)0;
    const T *ct = c4 + m0 * FF_C4_ROW;
    if constexpr (sizeof(T) == 8) {
#pragma unroll
        for (int j = 0; j < 18; j += 2)
#pragma unroll
            for (int m = 0; m < 6; m++) {
                const double2 c = *reinterpret_cast<const double2 *>(ct + m * FF_C4_ROW + j);
                out[m] = FastArith<T>::fma(u[j], c.x, out[m]);
                out[m] = FastArith<T>::fma(u[j + 1], c.y, out[m]);
            }
    } else {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
#pragma unroll
            for (int m = 0; m < 6; m++) {
                const float4 c = *reinterpret_cast<const float4 *>(ct + m * FF_C4_ROW + j);
                out[m] = FastArith<T>::fma(u[j], c.x, out[m]);
                out[m] = FastArith<T>::fma(u[j + 1], c.y, out[m]);
                out[m] = FastArith<T>::fma(u[j + 2], c.z, out[m]);
                out[m] = FastArith<T>::fma(u[j + 3], c.w, out[m]);
            }
#pragma unroll
        for (int m = 0; m < 6; m++) {
            const float2 c = *reinterpret_cast<const float2 *>(ct + m * FF_C4_ROW + 16);
            out[m] = FastArith<T>::fma(u[16], c.x, out[m]);
            out[m] = FastArith<T>::fma(u[17], c.y, out[m]);
        }
    }
}

// short window L of a band (mdct.c:171-185): 12 -> 6 through the same fold, 6-point DCT-IV from constant memory
template <int L, class T>
__device__ __forceinline__ void ff_mdct_short6(const T (&in)[36], T (&out)[6])
{
    T f[12], u[6];
#pragma unroll
    for (int k = 0; k < 12; k++) f[k] = FastArith<T>::mul(FastTab<T>::win(2, k), in[k + 6 * L + 6]);
#pragma unroll
    for (int j = 0; j < 3; j++) u[j] = FastArith<T>::sub(-f[8 - j], f[9 + j]);
#pragma unroll
    for (int j = 3; j < 6; j++) u[j] = FastArith<T>::sub(f[j - 3], f[8 - j]);
#pragma unroll
    for (int m = 0; m < 6; m++) {
        T s = (T)0;
#pragma unroll
        for (int j = 0; j < 6; j++) s = FastArith<T>::fma(u[j], FastTab<T>::dct4_s(m, j), s);
        out[m] = s;
    }
}

// ---- stage B on the FP64 tensor cores (the A/B BASELINE's north star asks for): S[32 slots][32 bands] = Y[32][32] x AM'^T with
// mma.sync.m8n8k4.f64 (SASS DMMA), one warp per chunk of 32 slots, eight slots (one m-tile) at a time.  AM'[sb][31] = 1
// carries y16 (the j = 31 column of the row layout), so the whole of encode.c:399-408 is the one product.  Fragments
// (PTX ISA, m8n8k4 .f64): lane T holds A[T/4][T%4], B[T%4][T/4], C[T/4][2 (T%4) + {0, 1}].
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void ff_matrix_dmma(double *rows /* the warp's 32 rows */, const double *amp /* [32][FT_ROW], [sb][31] = 1 */,
                                               int lane, int slot0)
{
    const int r = lane >> 2, q = lane & 3;
#pragma unroll 1
    for (int mt = 0; mt < 4; mt++) {
        double a[8], acc[4][2];
        const double *yrow = rows + (size_t)(8 * mt + r) * FT_ROW;
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] = yrow[4 * k + q];
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
            acc[nt][0] = 0.0; acc[nt][1] = 0.0;
            const double *brow = amp + (size_t)(8 * nt + r) * FT_ROW;        // B[k][n] = AM'[n][k]
#pragma unroll
            for (int k = 0; k < 8; k++) dmma884(acc[nt][0], acc[nt][1], a[k], brow[4 * k + q]);
        }
        __syncwarp();                                                        // every lane has read its part of the eight rows
        const bool odd_slot = (((slot0 + 8 * mt + r) % 18) & 1) != 0;
        double *orow = rows + (size_t)(8 * mt + r) * FT_ROW;
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
            const int sb = 8 * nt + 2 * q;                                   // even band, then the odd one: mdct.c:57-60
            orow[sb] = acc[nt][0];
            orow[sb + 1] = odd_slot ? -acc[nt][1] : acc[nt][1];
        }
        __syncwarp();
    }
}

template <class T, bool LEE, class OUT, bool TC = false>
__global__ void __launch_bounds__(FT_THREADS, (sizeof(T) == 4) ? 3 : 2)
k_front_fast(const short *__restrict__ pcm_rows, long stream_stride, long ch_stride, int hist, int n_streams, int n_ch, int n_gran,
             const int *__restrict__ nfr, const PsyOut *__restrict__ psy, OUT *__restrict__ xr)
{
    using A = FastArith<T>;
    extern __shared__ __align__(16) unsigned char ft_smem_raw[];
    FrontFastSmem<T, LEE> &M = *reinterpret_cast<FrontFastSmem<T, LEE> *>(ft_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_tiles = (n_gran + FT_G - 1) / FT_G;
    const long bid = blockIdx.x;
    const int t = (int)(bid % n_tiles);
    const int ch = (int)((bid / n_tiles) % n_ch);
    const long s = bid / ((long)n_tiles * n_ch);
    const int g_first = t * FT_G;
    const int n_live = nfr ? min(n_gran, 2 * nfr[s]) : n_gran;
    const int ng = min(FT_G, n_live - g_first);
    if (ng <= 0) return;
    const int n_slots = 18 * (ng + 1);
    const int n_chunks = (n_slots + 31) >> 5;
    // ---- stage 0: tables (converted to T) and the PCM tile --------------------------------------------------------------
    for (int i = tid; i < 512; i += FT_THREADS) M.window[i] = FastTab<T>::window(i);
    if constexpr (!LEE && !TC)
        for (int i = tid; i < 32 * 32; i += FT_THREADS) M.am[i] = (T)(&g_front.am[0][0])[i];
    if constexpr (TC)      // padded rows (conflict-free-ish fragment loads), [sb][31] = 1 for the y16 column
        for (int i = tid; i < 32 * FT_ROW; i += FT_THREADS) {
            const int sb = i / FT_ROW, j = i - sb * FT_ROW;
            M.am[i] = j < 31 ? (T)g_front.am[sb][j] : (j == 31 ? (T)1 : (T)0);
        }
    for (int i = tid; i < 18 * FF_C4_ROW; i += FT_THREADS) {
        const int m = i / FF_C4_ROW, j = i - m * FF_C4_ROW;
        M.c4[i] = j < 18 ? FastTab<T>::dct4_l(m, j) : (T)0;
    }
    {
        const short *src = pcm_rows + s * stream_stride + ch * ch_stride + hist + 576L * (g_first - 1) - 480;
        const int n_valid = 480 + 32 * n_slots;
        const uint4 *src4 = reinterpret_cast<const uint4 *>(src);
        uint4 *dst4 = reinterpret_cast<uint4 *>(M.pcm);
        for (int i = tid; i < FT_PCM / 8; i += FT_THREADS) dst4[i] = (i < n_valid / 8) ? src4[i] : make_uint4(0, 0, 0, 0);
        if (tid < ng) M.bt[tid] = psy[((s * n_gran + g_first + tid) * (long)n_ch + ch)].block_type;
    }
    __syncthreads();
    // ---- stage A: lane = tap i (and i + 32); slots of one parity form a sliding 8-tap FIR (encode.c:306-311, 392-398) -----
    for (int c = warp; c < n_chunks; c += FT_WARPS) {
        T w0[8], w1[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { w0[j] = M.window[lane + 64 * j]; w1[j] = M.window[lane + 32 + 64 * j]; }
        const int src_lane = (32 - lane) & 31;
        const T scale = (T)(1.0 / 32768);
#pragma unroll 1
        for (int p = 0; p < 2; p++) {
            const int base0 = 480 + 32 * (32 * c + p) + 31 - lane;
            T h0[8], h1[8];
#pragma unroll
            for (int j = 1; j < 8; j++) {
                h0[j] = A::mul((T)M.pcm[base0 - 64 * j], scale);
                h1[j] = A::mul((T)M.pcm[base0 - 32 - 64 * j], scale);
            }
#pragma unroll
            for (int k = 0; k < 16; k++) {
                h0[0] = A::mul((T)M.pcm[base0 + 64 * k], scale);
                h1[0] = A::mul((T)M.pcm[base0 - 32 + 64 * k], scale);
                T y0 = A::mul(h0[0], w0[0]), y1 = A::mul(h1[0], w1[0]);
#pragma unroll
                for (int j = 1; j < 8; j++) { y0 = A::fma(h0[j], w0[j], y0); y1 = A::fma(h1[j], w1[j], y1); }
#pragma unroll
                for (int j = 7; j > 0; j--) { h0[j] = h0[j - 1]; h1[j] = h1[j - 1]; }
                const T a0 = __shfl_sync(0xffffffffu, y0, src_lane);               // y[32 - i]
                const T a1 = __shfl_sync(0xffffffffu, y1, src_lane);               // y[64 - i]
                T *row = M.rows + (size_t)(32 * c + 2 * k + p) * FT_ROW;
                if constexpr (LEE) {
                    // Lee's DCT-III input order: t[0] = y[16], t[n] = ysum[16 - n] (n = 1..16), t[n] = ysub[n - 17] (n = 17..31)
                    if (lane == 0) row[16] = A::add(y0, y1);                        // ysum[0] = y[0] + y[32]
                    else if (lane < 16) {
                        row[16 - lane] = A::add(y0, a0);                            // ysum[i] = y[i] + y[32 - i]
                        row[16 + lane] = A::sub(y1, a1);                            // ysub[i - 1] = y[32 + i] - y[64 - i]
                    } else if (lane == 16) row[0] = y0;                             // y[16]
                } else {
                    if (lane == 0) row[0] = A::add(y0, y1);
                    else if (lane < 16) {
                        row[lane] = A::add(y0, a0);
                        row[15 + lane] = A::sub(y1, a1);
                    } else if (lane == 16) row[31] = y0;
                }
            }
        }
    }
    __syncwarp();
    // ---- stage B: thread = slot (or, on the tensor cores, warp = chunk of 32 slots) -----------------------------------------
    if constexpr (TC) {
        if (warp < n_chunks) ff_matrix_dmma(M.rows + (size_t)(32 * warp) * FT_ROW, M.am, lane, 32 * warp);
    } else if (tid < n_chunks * 32) {
        T *row = M.rows + (size_t)tid * FT_ROW;
        T ys[32];
#pragma unroll
        for (int j = 0; j < 32; j++) ys[j] = row[j];
        const bool odd_slot = ((tid % 18) & 1) != 0;
        if constexpr (LEE) {
            lee_dct3<32, T>(ys);
#pragma unroll
            for (int j = 0; j < 32; j++) row[j] = ((j & 1) && odd_slot) ? -ys[j] : ys[j];   // mdct.c:57-60
        } else {
#pragma unroll 1
            for (int sb0 = 0; sb0 < 32; sb0 += 8) ff_matrix_rows8<T>(ys, row, sb0, odd_slot, M.am);
        }
    }
    __syncthreads();
    // ---- stage C + D: (granule, third) per warp, lane = band ----------------------------------------------------------------
    const long gc_stride = n_ch;
    const int third = warp % 3, group = warp / 3;
    for (int r = 0; 3 * r < ng; r++) {
        const int gl = 1 + 3 * r + group;
        const bool active = gl <= ng;
        T in[36], out[6];
        int bt = 0;
        if (active) {
            const T *p = M.rows + (size_t)(18 * (gl - 1)) * FT_ROW + lane;
#pragma unroll
            for (int k = 0; k < 36; k++) in[k] = p[k * FT_ROW];
            bt = M.bt[gl - 1];
        }
        __syncthreads();
        if (active) {
            T *st = M.rows + (size_t)(18 * (gl - 1)) * FT_ROW;
            if (bt == 2) {
                if (third == 0) ff_mdct_short6<0, T>(in, out); else if (third == 1) ff_mdct_short6<1, T>(in, out); else ff_mdct_short6<2, T>(in, out);
#pragma unroll
                for (int m = 0; m < 6; m++) st[lane * 18 + 3 * m + third] = out[m];
            } else {
                T u[18];
                if (bt == 0) ff_fold_long<0, T>(in, u); else if (bt == 1) ff_fold_long<1, T>(in, u); else ff_fold_long<3, T>(in, u);
#if FF_COEF_CONST
                if (third == 0) ff_dct4_6c<0, T>(u, out); else if (third == 1) ff_dct4_6c<6, T>(u, out); else ff_dct4_6c<12, T>(u, out);
#else
                ff_dct4_6<T>(u, out, 6 * third, M.c4);
#endif
#pragma unroll
                for (int m = 0; m < 6; m++) st[lane * 18 + 6 * third + m] = out[m];
            }
            ft_group_barrier(group);
            if (bt != 2 && lane < 31) {                                                   // mdct.c:83-91
                for (int k = third; k < 8; k += 3) {
                    const T a = st[lane * 18 + 17 - k], b = st[(lane + 1) * 18 + k];
                    const T cs = FastTab<T>::cs(k), ca = FastTab<T>::ca(k);
                    st[lane * 18 + 17 - k] = A::fma(b, ca, A::mul(a, cs));
                    st[(lane + 1) * 18 + k] = A::fma(-a, ca, A::mul(b, cs));
                }
            }
            ft_group_barrier(group);
            OUT *dst = xr + ((s * n_gran + g_first + gl - 1) * gc_stride + ch) * 576;
            if constexpr (sizeof(T) == sizeof(OUT) && sizeof(T) == 8) {
                double2 *d2 = reinterpret_cast<double2 *>(dst);
                const double2 *s2 = reinterpret_cast<const double2 *>(st);
#pragma unroll
                for (int i = 0; i < 3; i++) d2[lane + 32 * (3 * third + i)] = s2[lane + 32 * (3 * third + i)];
            } else if constexpr (sizeof(T) == sizeof(OUT)) {
                // 576 floats: 144 float4, 48 per third; 16-byte alignment of st: (18 (gl - 1) FT_ROW) floats — FT_ROW is odd, so
                // fall back to scalar-pair loads from shared memory and 8-byte coalesced stores
                float2 *d2 = reinterpret_cast<float2 *>(dst);
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    const int e = lane + 32 * (3 * third + i);
                    d2[e] = make_float2(st[2 * e], st[2 * e + 1]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 6; i++) {
                    const int e = lane + 32 * (6 * third + i);
                    dst[e] = (OUT)st[e];
                }
            }
        }
    }
}

// ---- FP32 variant, second version (round 2): same dataflow and arithmetic as k_front_fast<float, true, float>, restructured
// around what ncu showed (profiles/r02_front_variants.md): the PCM tile arrives by one bulk copy (TMA) while the threads
// fetch their window taps straight from global memory (no table staging, no barrier for it); the 1/32768 scale rides on
// the taps; the fold writes ONE value per lane (lanes 1..15 the sums, 17..31 the differences, one shuffle); the MDCT of a
// granule is ONE warp's work (window products and the aliasing fold once instead of three times, a third of the input
// loads, no barriers among warps of a granule), its 18 x 18 DCT-IV entirely on constant-bank operands.
struct FrontF32Smem {
    float rows[FT_SLOTS * FT_ROW];
    short pcm[FT_PCM];
    int bt[FT_G + 1];
    unsigned long long bar;
};

// 18-point DCT-IV on the folded input, X[m] = sum_j u[j] cos(pi/72 (2j+1)(2m+1)) / 9, through ONE 9-point complex DFT:
//   z[n] = u[2n] + i u[17 - 2n];  t[n] = z[n] e^{-i pi n / 18};  T = DFT9(t);  c[k] = T[k] e^{-i pi (4k+1) / 72} / 9;
//   X[2k] = Re c[k],  X[17 - 2k] = -Im c[k]
// (split the sum over even j and over j = 17 - 2n: the cosine of the second half is the sine of the first half's angle), the
// DFT9 as 3 x 3 (two rounds of 3-point transforms, four twiddles between them).  156 fused operations on immediates instead of
// 324 on loaded coefficients; the arithmetic differs from the direct form by FP32 rounding only (a few 1e-7 of the largest
// value: the tolerance of this path is 1e-5).
struct FfC { float r, i; };
__device__ __forceinline__ void ff_dft3(FfC &x0, FfC &x1, FfC &x2)
{
    const float h = 0.866025404f;
    const float sr = x1.r + x2.r, si = x1.i + x2.i, dr = x1.r - x2.r, di = x1.i - x2.i;
    const float mr = __fmaf_rn(-0.5f, sr, x0.r), mi = __fmaf_rn(-0.5f, si, x0.i);
    x0.r += sr; x0.i += si;
    x1.r = __fmaf_rn(h, di, mr); x1.i = __fmaf_rn(-h, dr, mi);
    x2.r = __fmaf_rn(-h, di, mr); x2.i = __fmaf_rn(h, dr, mi);
}
__device__ __forceinline__ FfC ff_rot(FfC x, float c, float s)     // x * (c - i s)
{
    FfC y;
    y.r = __fmaf_rn(x.i, s, x.r * c); y.i = __fmaf_rn(-x.r, s, x.i * c);
    return y;
}
__device__ __forceinline__ void ff_dct4_18(const float (&u)[18], float (&out)[18])
{
    constexpr float pre_c[9] = {1.f, 0.984807753f, 0.939692621f, 0.866025404f, 0.766044443f, 0.64278761f, 0.5f, 0.342020143f, 0.173648178f};
    constexpr float pre_s[9] = {0.f, 0.173648178f, 0.342020143f, 0.5f, 0.64278761f, 0.766044443f, 0.866025404f, 0.939692621f, 0.984807753f};
    constexpr float post_c[9] = {0.111005358f, 0.108477334f, 0.102653281f, 0.0937101606f, 0.0819197041f, 0.0676401588f, 0.0513054015f, 0.0334117555f, 0.0145029102f};
    constexpr float post_s[9] = {0.0048465986f, 0.024048846f, 0.0425203814f, 0.0596999565f, 0.0750655786f, 0.0881503711f, 0.0985567592f, 0.10596855f, 0.11016054f};
    FfC t[9];
    t[0].r = u[0]; t[0].i = u[17];
#pragma unroll
    for (int n = 1; n < 9; n++) { FfC z; z.r = u[2 * n]; z.i = u[17 - 2 * n]; t[n] = ff_rot(z, pre_c[n], pre_s[n]); }
    // DFT9, n = 3a + b, k = k1 + 3 k2: three DFT3 over a, twiddles W9^(b k1), three DFT3 over b
#pragma unroll
    for (int b = 0; b < 3; b++) ff_dft3(t[b], t[3 + b], t[6 + b]);            // t[3 k1 + b] = A[b][k1]
    t[4] = ff_rot(t[4], 0.766044443f, 0.64278761f);                           // b = 1, k1 = 1: W9^1
    t[7] = ff_rot(t[7], 0.173648178f, 0.984807753f);                          // b = 1, k1 = 2: W9^2
    t[5] = ff_rot(t[5], 0.173648178f, 0.984807753f);                          // b = 2, k1 = 1: W9^2
    t[8] = ff_rot(t[8], -0.939692621f, 0.342020143f);                         // b = 2, k1 = 2: W9^4
#pragma unroll
    for (int k1 = 0; k1 < 3; k1++) ff_dft3(t[3 * k1], t[3 * k1 + 1], t[3 * k1 + 2]);   // t[3 k1 + k2] = T[k1 + 3 k2]
#pragma unroll
    for (int k1 = 0; k1 < 3; k1++)
#pragma unroll
        for (int k2 = 0; k2 < 3; k2++) {
            const int k = k1 + 3 * k2;
            const FfC T = t[3 * k1 + k2];
            out[2 * k] = __fmaf_rn(T.i, post_s[k], T.r * post_c[k]);
            out[17 - 2 * k] = __fmaf_rn(-T.i, post_c[k], T.r * post_s[k]);
        }
}

// A persistent version (3 CTAs per SM walking the tiles, the next tile's PCM fetched into the same buffer behind the MDCT
// stage) was measured slower: 16.3 against 12.9 ms per 4144-clip step (the loop state spills under the 72-register cap and
// costs two more CTA barriers per tile) — as FT_PERSISTENT was for the exact kernel.
__global__ void __launch_bounds__(FT_THREADS, 3)
k_front_f32(const short *__restrict__ pcm_rows, long stream_stride, long ch_stride, int hist, int n_streams, int n_ch, int n_gran,
            const int *__restrict__ nfr, const PsyOut *__restrict__ psy, float *__restrict__ xr)
{
    using A = FastArith<float>;
    extern __shared__ __align__(16) unsigned char ft_smem_raw[];
    FrontF32Smem &M = *reinterpret_cast<FrontF32Smem *>(ft_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_tiles = (n_gran + FT_G - 1) / FT_G;
    const long bid = blockIdx.x;
    const int t = (int)(bid % n_tiles);
    const int ch = (int)((bid / n_tiles) % n_ch);
    const long s = bid / ((long)n_tiles * n_ch);
    const int g_first = t * FT_G;
    const int n_live = nfr ? min(n_gran, 2 * nfr[s]) : n_gran;
    const int ng = min(FT_G, n_live - g_first);
    if (ng <= 0) return;
    const int n_slots = 18 * (ng + 1);
    const int n_chunks = (n_slots + 31) >> 5;
    const int n_valid = 480 + 32 * n_slots;               // multiple of 8 samples = 16 bytes
    // ---- stage 0 ----------------------------------------------------------------------------------------------------------
    if (tid == 0) {
        ft_mbar_init(&M.bar, 1);
        const short *src = pcm_rows + s * stream_stride + ch * ch_stride + hist + 576L * (g_first - 1) - 480;
        ft_bulk_load(M.pcm, src, (unsigned)(n_valid * sizeof(short)), &M.bar);
    }
    {
        uint4 *dst4 = reinterpret_cast<uint4 *>(M.pcm);
        for (int i = n_valid / 8 + tid; i < FT_PCM / 8; i += FT_THREADS) dst4[i] = make_uint4(0, 0, 0, 0);   // a short last tile
        if (tid < ng) M.bt[tid] = psy[((s * n_gran + g_first + tid) * (long)n_ch + ch)].block_type;
    }
    float w0[8], w1[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {       // encode.c:306: /SCALE, a power of two, moved onto the taps
        w0[j] = A::mul(g_front_f.window[lane + 64 * j], 1.0f / 32768);
        w1[j] = A::mul(g_front_f.window[lane + 32 + 64 * j], 1.0f / 32768);
    }
    // the fold as one fused expression per lane: v = f0 y0 + f1 y1 + fo o  (lane 0: y0 + y1; 1..15: y0 + o; 16: y0; 17..31: o - y1)
    const float f0 = lane <= 16 ? 1.f : 0.f, f1 = lane == 0 ? 1.f : lane > 16 ? -1.f : 0.f, fo = (lane == 0 || lane == 16) ? 0.f : 1.f;
    const int src_lane = (32 - lane) & 31;
    // Lee's DCT-III input order: t[0] = y[16], t[n] = ysum[16 - n] (n = 1..16), t[n] = ysub[n - 17] (n = 17..31)
    const int idx = lane == 0 ? 16 : lane <= 16 ? 16 - lane : 48 - lane;
    __syncthreads();                                      // the mbarrier is initialised; bt[] and the zero fill are visible
    ft_mbar_wait(&M.bar, 0);
    // ---- stage A: lane = tap i (and i + 32); slots of one parity form a sliding 8-tap FIR (encode.c:306-311, 392-398) -----
    for (int c = warp; c < n_chunks; c += FT_WARPS) {
#pragma unroll 1
        for (int p = 0; p < 2; p++) {
            const int base0 = 480 + 32 * (32 * c + p) + 31 - lane;
            float h0[8], h1[8];
#pragma unroll
            for (int j = 1; j < 8; j++) { h0[j] = (float)M.pcm[base0 - 64 * j]; h1[j] = (float)M.pcm[base0 - 32 - 64 * j]; }
#pragma unroll
            for (int k = 0; k < 16; k++) {
                h0[0] = (float)M.pcm[base0 + 64 * k];
                h1[0] = (float)M.pcm[base0 - 32 + 64 * k];
                float y0 = A::mul(h0[0], w0[0]), y1 = A::mul(h1[0], w1[0]);
#pragma unroll
                for (int j = 1; j < 8; j++) { y0 = A::fma(h0[j], w0[j], y0); y1 = A::fma(h1[j], w1[j], y1); }
#pragma unroll
                for (int j = 7; j > 0; j--) { h0[j] = h0[j - 1]; h1[j] = h1[j - 1]; }
                // lanes 1..15 need y[32 - i] (y0 of lane 32 - i), lanes 17..31 need y[32 + (32 - lane)] (y1 of lane 32 - lane)
                const float o = __shfl_sync(0xffffffffu, lane < 16 ? y1 : y0, src_lane);
                M.rows[(size_t)(32 * c + 2 * k + p) * FT_ROW + idx] = A::fma(fo, o, A::fma(f1, y1, A::mul(f0, y0)));
            }
        }
    }
    __syncwarp();
    // ---- stage B: thread = slot, Lee's DCT-III in registers ----------------------------------------------------------------
    if (tid < n_chunks * 32) {
        float *row = M.rows + (size_t)tid * FT_ROW;
        float ys[32];
#pragma unroll
        for (int j = 0; j < 32; j++) ys[j] = row[j];
        const bool odd_slot = ((tid % 18) & 1) != 0;
        lee_dct3<32, float>(ys);
#pragma unroll
        for (int j = 0; j < 32; j++) row[j] = ((j & 1) && odd_slot) ? -ys[j] : ys[j];   // mdct.c:57-60
    }
    __syncthreads();
    // ---- stage C + D: one warp per granule, lane = band --------------------------------------------------------------------
    for (int r = 0; FT_WARPS * r < ng; r++) {
        const int gl = 1 + FT_WARPS * r + warp;
        const bool active = gl <= ng;
        float in[36];
        int bt = 0;
        if (active) {
            const float *p = M.rows + (size_t)(18 * (gl - 1)) * FT_ROW + lane;
#pragma unroll
            for (int k = 0; k < 36; k++) in[k] = p[k * FT_ROW];
            bt = M.bt[gl - 1];
        }
        __syncthreads();                                  // every warp has its inputs: the rows of the previous granules are free
        if (active) {
            float *st = M.rows + (size_t)(18 * (gl - 1)) * FT_ROW;
            if (bt == 2) {
                float out[6];
                ff_mdct_short6<0, float>(in, out);
#pragma unroll
                for (int m = 0; m < 6; m++) st[lane * 18 + 3 * m] = out[m];
                ff_mdct_short6<1, float>(in, out);
#pragma unroll
                for (int m = 0; m < 6; m++) st[lane * 18 + 3 * m + 1] = out[m];
                ff_mdct_short6<2, float>(in, out);
#pragma unroll
                for (int m = 0; m < 6; m++) st[lane * 18 + 3 * m + 2] = out[m];
            } else {
                float u[18], out[18];
                if (bt == 0) ff_fold_long<0, float>(in, u); else if (bt == 1) ff_fold_long<1, float>(in, u); else ff_fold_long<3, float>(in, u);
                ff_dct4_18(u, out);
#pragma unroll
                for (int m = 0; m < 18; m++) st[lane * 18 + m] = out[m];
            }
            __syncwarp();
            if (bt != 2 && lane < 31) {                                                   // mdct.c:83-91
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const float a = st[lane * 18 + 17 - k], b = st[(lane + 1) * 18 + k];
                    const float cs = FastTab<float>::cs(k), ca = FastTab<float>::ca(k);
                    st[lane * 18 + 17 - k] = A::fma(b, ca, A::mul(a, cs));
                    st[(lane + 1) * 18 + k] = A::fma(-a, ca, A::mul(b, cs));
                }
            }
            __syncwarp();
            float2 *d2 = reinterpret_cast<float2 *>(xr + ((s * n_gran + g_first + gl - 1) * (long)n_ch + ch) * 576);
#pragma unroll
            for (int i = 0; i < 9; i++) {
                const int e = lane + 32 * i;
                d2[e] = make_float2(st[2 * e], st[2 * e + 1]);
            }
        }
    }
}

}  // namespace mp3gpu
