// fft_regs.h — the psychoacoustic model's real FFTs (rsfft / rsrec / srrec, /root/reference/src/subs.c:185-534) as
// straight-line REGISTER code of one warp: no op interpreter, no shared-memory round trip per butterfly.
//
// The reference's recursive split-radix transform is a fixed dataflow graph; FP32 results are bit-identical as long as
// every add / multiply keeps its operands (IEEE add and multiply are commutative; x - y == x + (-y); rounding is
// symmetric in the sign).  Mapping (tables.h, "FFT in registers"):
//   phase 1  element j in register j / 32 of lane j % 32: the steps of all nodes with strides >= 32, same code in all lanes
//   phase 2  one 32-element block per lane (through shared memory, blocks sorted by kind over two passes): the nodes of
//            length <= 32 and the stride-16 steps of the length-64 complex nodes; the real and the imaginary array of a
//            complex node are two lanes that exchange operands by shuffle
// Sign changes that the reference applies before further arithmetic (rsrec step 2) are applied at once; those that
// follow the last arithmetic on a value (rsrec step 5, `if (logm == 2) x[3] = -x[3]`) and all data reorders (step 5,
// BR_permute) are folded into the output maps, exactly as tables.cpp's build_fft_program tracks them.
//
// Per-lane code: compiled by nvcc for the device, and by g++ (-DMP3GPU_HOST_EMUL) for tests/emul, where the 32 lanes run
// as fibers and lanes::shfl / lanes::sync are rendezvous points.
#pragma once
#include "simt.h"
#include "tables.h"

namespace mp3gpu {

#if SIMT_DEV
__constant__ FftRegsConst c_fftr;   // defined here: this header is included by exactly one translation unit (mp3gpu.cu)
namespace lanes {
SIMT_FN float shfl(unsigned mask, float v, int src) { return __shfl_sync(mask, v, src); }
SIMT_FN void sync() { __syncwarp(); }
SIMT_FN float fxor(float v, unsigned s) { return __uint_as_float(__float_as_uint(v) ^ s); }
}  // namespace lanes
struct FftVec4 { float x, y, z, w; };
SIMT_FN FftVec4 fftr_ld4(const float *p) { const float4 v = *reinterpret_cast<const float4 *>(p); return {v.x, v.y, v.z, v.w}; }
SIMT_FN void fftr_st4(float *p, float a, float b, float c, float d) { *reinterpret_cast<float4 *>(p) = make_float4(a, b, c, d); }
SIMT_FN FftVec4 fftr_ldg4(const float *p) { const float4 v = __ldg(reinterpret_cast<const float4 *>(p)); return {v.x, v.y, v.z, v.w}; }
#else
extern FftRegsConst c_fftr;         // tests/emul
namespace lanes {
float shfl(unsigned mask, float v, int src);   // fiber rendezvous of the lanes in `mask` (tests/emul/emul.cpp)
void sync();
inline float fxor(float v, unsigned s) { unsigned u; std::memcpy(&u, &v, 4); u ^= s; std::memcpy(&v, &u, 4); return v; }
}  // namespace lanes
struct FftVec4 { float x, y, z, w; };
inline FftVec4 fftr_ld4(const float *p) { return {p[0], p[1], p[2], p[3]}; }
inline void fftr_st4(float *p, float a, float b, float c, float d) { p[0] = a; p[1] = b; p[2] = c; p[3] = d; }
inline FftVec4 fftr_ldg4(const float *p) { return {p[0], p[1], p[2], p[3]}; }
#endif

namespace fftr {

using simt::fadd;
using simt::fmul;
using simt::fsub;

SIMT_FN float sqmul(float t) { return (float)simt::dmul(0.707106781186547524401, (double)t); }   // SQHALF * t, subs.c:26 (double multiply)
SIMT_FN void bfly(float &a, float &b) { const float t = fadd(a, b); b = fsub(a, b); a = t; }     // t=a+b; b=a-b; a=t
template <int M> struct ILog2 { static constexpr int v = 1 + ILog2<M / 2>::v; };
template <> struct ILog2<1> { static constexpr int v = 0; };

template <int L>
SIMT_FN FftTwC small_tw(int set, int n)
{
    if constexpr (L == 4) return c_fftr.small.t4[set][n];
    else if constexpr (L == 5) return c_fftr.small.t5[set][n];
    else return c_fftr.small.t6[set][n];
}

// ---- in-lane nodes: real and imaginary array in the same lane ----------------------------------------------------------
SIMT_FN void cross(float &r1, float &r2, float &i1, float &i2)      // srrec step 2, subs.c:293-304
{
    const float t1 = fadd(r1, i2), t2 = fadd(i1, r2);
    i1 = fsub(i1, r2); r2 = fsub(r1, i2); r1 = t1; i2 = t2;
}
SIMT_FN void rot(float &a, float &c, const FftTwC &T)               // subs.c:329-337 / 486-490
{
    const float t2 = fmul(T.cn, fadd(a, c)), t1 = fadd(fmul(T.spcn, a), t2);
    a = fadd(fmul(T.smcn, c), t2); c = t1;
}
SIMT_FN void rot8a(float &a, float &c) { const float t1 = sqmul(fadd(a, c)); c = sqmul(fsub(c, a)); a = t1; }        // subs.c:321-323
SIMT_FN void rot8b(float &b, float &d) { const float t2 = sqmul(fsub(d, b)); d = -sqmul(fadd(b, d)); b = t2; }       // subs.c:324-326

template <int M, int RU, int IU>     // steps 2-4 of a complex node of length M; RU / IU: registers of elements M/2 .. M of xr / xi
SIMT_FN void sr_steps234(float (&x)[32])
{
    constexpr int m4 = M / 4, m8 = M / 8, L = ILog2<M>::v;
#pragma unroll
    for (int n = 0; n < m4; n++) cross(x[RU + n], x[RU + m4 + n], x[IU + n], x[IU + m4 + n]);
#pragma unroll
    for (int n = 1; n < m4; n++) {
        if (n == m8) { rot8a(x[RU + n], x[IU + n]); rot8b(x[RU + m4 + n], x[IU + m4 + n]); }
        else if constexpr (L >= 4) { rot(x[RU + n], x[IU + n], small_tw<L>(0, n)); rot(x[RU + m4 + n], x[IU + m4 + n], small_tw<L>(1, n)); }
    }
}

template <int M, int RR, int RI>     // complex node of length M, xr in registers RR.., xi in registers RI..
SIMT_FN void sr_full(float (&x)[32])
{
    if constexpr (M == 2) { bfly(x[RR], x[RR + 1]); bfly(x[RI], x[RI + 1]); }
    else if constexpr (M == 4) {     // subs.c:202-238
        bfly(x[RR], x[RR + 2]); bfly(x[RI], x[RI + 2]); bfly(x[RR + 1], x[RR + 3]); bfly(x[RI + 1], x[RI + 3]);
        bfly(x[RR], x[RR + 1]); bfly(x[RI], x[RI + 1]);
        cross(x[RR + 2], x[RR + 3], x[RI + 2], x[RI + 3]);
    } else if constexpr (M >= 8) {
        constexpr int m2 = M / 2, m4 = M / 4;
#pragma unroll
        for (int n = 0; n < m2; n++) { bfly(x[RR + n], x[RR + n + m2]); bfly(x[RI + n], x[RI + n + m2]); }
        sr_steps234<M, RR + m2, RI + m2>(x);
        sr_full<M / 2, RR, RI>(x);
        sr_full<M / 4, RR + m2, RI + m2>(x);
        sr_full<M / 4, RR + 3 * m4, RI + 3 * m4>(x);
    }
}

template <int M, int RU>             // steps 2-4 of a real node of length M; RU: register of element M/2
SIMT_FN void rs_steps234(float (&x)[32])
{
    constexpr int m4 = M / 4, m8 = M / 8, L = ILog2<M>::v;
#pragma unroll
    for (int n = 0; n < m4; n++) x[RU + m4 + n] = -x[RU + m4 + n];
#pragma unroll
    for (int n = 1; n < m4; n++) {
        if (n == m8) rot8a(x[RU + n], x[RU + m4 + n]);
        else if constexpr (L >= 4) rot(x[RU + n], x[RU + m4 + n], small_tw<L>(0, n));
    }
}

template <int M, int R0>             // real node of length M on registers R0 .. R0 + M
SIMT_FN void rs_full(float (&x)[32])
{
    if constexpr (M == 2) bfly(x[R0], x[R0 + 1]);
    else if constexpr (M == 4) {     // the sign change of step 2 and `if (logm == 2) x[3] = -x[3]` cancel
        bfly(x[R0], x[R0 + 2]); bfly(x[R0 + 1], x[R0 + 3]); bfly(x[R0], x[R0 + 1]);
    } else if constexpr (M >= 8) {
        constexpr int m2 = M / 2, m4 = M / 4;
#pragma unroll
        for (int n = 0; n < m2; n++) bfly(x[R0 + n], x[R0 + n + m2]);
        rs_steps234<M, R0 + m2>(x);
        rs_full<M / 2, R0>(x);
        sr_full<M / 4, R0 + m2, R0 + 3 * m4>(x);
    }
}

// ---- half nodes: this lane holds ONE of the two arrays of a complex node, lane `partner` the other ----------------------
struct Half {
    unsigned mask;      // lanes that run this code together
    int partner;
    unsigned sgn;       // 0x80000000 in the lane that holds the imaginary array, else 0
    bool is_xi;
};

// srrec step 2: xr lane: (xr1, xr2) <- (xr1 + xi2, xr1 - xi2);  xi lane: (xi1, xi2) <- (xi1 - xr2, xi1 + xr2)
SIMT_FN void cross_half(float &x1, float &x2, const Half &h)
{
    const float v = lanes::fxor(lanes::shfl(h.mask, x2, h.partner), h.sgn), t = x1;
    x1 = fadd(t, v); x2 = fsub(t, v);
}
// tmp2 = cn (xr + xi); xr lane: xr <- smcn xi + tmp2;  xi lane: xi <- spcn xr + tmp2
SIMT_FN void rot_half(float &mine, const FftTwC &T, const Half &h)
{
    const float v = lanes::shfl(h.mask, mine, h.partner);
    const float t2 = fmul(T.cn, fadd(mine, v));
    mine = fadd(fmul(h.is_xi ? T.spcn : T.smcn, v), t2);
}
// xr lane: xr1 <- SQ (xr1 + xi1);  xi lane: xi1 <- SQ (xi1 - xr1)
SIMT_FN void rot8a_half(float &mine, const Half &h)
{
    const float v = lanes::shfl(h.mask, mine, h.partner);
    mine = sqmul(fadd(mine, lanes::fxor(v, h.sgn)));
}
// xr lane: xr2 <- SQ (xi2 - xr2);  xi lane: xi2 <- -SQ (xr2 + xi2)
SIMT_FN void rot8b_half(float &mine, const Half &h)
{
    const float v = lanes::shfl(h.mask, mine, h.partner);
    mine = lanes::fxor(sqmul(fadd(v, lanes::fxor(mine, h.sgn ^ 0x80000000u))), h.sgn);
}

template <int M, int RU>             // steps 2-4 of a complex node of length M on this lane's array; RU: register of element M/2
SIMT_FN void sr_half_steps234(float (&x)[32], const Half &h)
{
    constexpr int m4 = M / 4, m8 = M / 8, L = ILog2<M>::v;
#pragma unroll
    for (int n = 0; n < m4; n++) cross_half(x[RU + n], x[RU + m4 + n], h);
#pragma unroll
    for (int n = 1; n < m4; n++) {
        if (n == m8) { rot8a_half(x[RU + n], h); rot8b_half(x[RU + m4 + n], h); }
        else if constexpr (L >= 4) { rot_half(x[RU + n], small_tw<L>(0, n), h); rot_half(x[RU + m4 + n], small_tw<L>(1, n), h); }
    }
}

template <int M, int R0>
SIMT_FN void sr_half(float (&x)[32], const Half &h)
{
    if constexpr (M == 2) bfly(x[R0], x[R0 + 1]);
    else if constexpr (M == 4) {
        bfly(x[R0], x[R0 + 2]); bfly(x[R0 + 1], x[R0 + 3]); bfly(x[R0], x[R0 + 1]);
        cross_half(x[R0 + 2], x[R0 + 3], h);
    } else if constexpr (M >= 8) {
        constexpr int m2 = M / 2, m4 = M / 4;
#pragma unroll
        for (int n = 0; n < m2; n++) bfly(x[R0 + n], x[R0 + n + m2]);
        sr_half_steps234<M, R0 + m2>(x, h);
        sr_half<M / 2, R0>(x, h);
        sr_half<M / 4, R0 + m2>(x, h);
        sr_half<M / 4, R0 + 3 * m4>(x, h);
    }
}

// ---- phase 1: element j = lane + 32 r in register r; N = registers of the whole array --------------------------------
// rotation of line n = lane + 32 K of a node of length 2^L (twiddle set SET); n == 0 is not rotated, n == m/8 is the
// SQHALF case (rot8a on the (xr1, xi1) / real pair = set 0, rot8b on the (xr2, xi2) pair = set 1)
template <int L, int SET, int K>
SIMT_FN void rotA(float &a, float &c, const float *twA, int lane)
{
    constexpr int m8 = (1 << L) / 8;
    const FftVec4 T = fftr_ldg4(twA + 4 * (FFTR_TWA_OFF(L, SET) + 32 * K + lane));
    const float t2 = fmul(T.x, fadd(a, c)), t1 = fadd(fmul(T.y, a), t2);
    float na = fadd(fmul(T.z, c), t2), nc = t1;
    if (K == 0 && lane == 0) { na = a; nc = c; }
    if (K == m8 / 32 && lane == m8 % 32) {
        na = a; nc = c;
        if (SET == 0) rot8a(na, nc); else rot8b(na, nc);
    }
    a = na; c = nc;
}

template <int L, int RR, int RI, int N>
SIMT_FN void srA(float (&x)[N], const float *twA, int lane)
{
    constexpr int NR = (1 << L) / 32;     // registers per component array
    if constexpr (L == 6) { bfly(x[RR], x[RR + 1]); bfly(x[RI], x[RI + 1]); }
    else if constexpr (L >= 7) {
        constexpr int h = NR / 2, q = NR / 4;
#pragma unroll
        for (int r = 0; r < h; r++) { bfly(x[RR + r], x[RR + r + h]); bfly(x[RI + r], x[RI + r + h]); }
#pragma unroll
        for (int k = 0; k < q; k++) cross(x[RR + h + k], x[RR + h + q + k], x[RI + h + k], x[RI + h + q + k]);
#pragma unroll
        for (int k = 0; k < q; k++) {
            if (k == 0) { rotA<L, 0, 0>(x[RR + h], x[RI + h], twA, lane); rotA<L, 1, 0>(x[RR + h + q], x[RI + h + q], twA, lane); }
            else if (k == NR / 8 && NR >= 8) { rotA<L, 0, (NR >= 8 ? NR / 8 : 0)>(x[RR + h + k], x[RI + h + k], twA, lane); rotA<L, 1, (NR >= 8 ? NR / 8 : 0)>(x[RR + h + q + k], x[RI + h + q + k], twA, lane); }
            else {
                const FftVec4 T0 = fftr_ldg4(twA + 4 * (FFTR_TWA_OFF(L, 0) + 32 * k + lane)), T1 = fftr_ldg4(twA + 4 * (FFTR_TWA_OFF(L, 1) + 32 * k + lane));
                rot(x[RR + h + k], x[RI + h + k], FftTwC{T0.x, T0.y, T0.z});
                rot(x[RR + h + q + k], x[RI + h + q + k], FftTwC{T1.x, T1.y, T1.z});
            }
        }
        srA<L - 1, RR, RI, N>(x, twA, lane);
        srA<L - 2, RR + h, RI + h, N>(x, twA, lane);
        srA<L - 2, RR + h + q, RI + h + q, N>(x, twA, lane);
    }
}

template <int L, int R0, int N>
SIMT_FN void rsA(float (&x)[N], const float *twA, int lane)
{
    constexpr int NR = (1 << L) / 32;
    if constexpr (L == 6) bfly(x[R0], x[R0 + 1]);
    else if constexpr (L >= 7) {
        constexpr int h = NR / 2, q = NR / 4;
#pragma unroll
        for (int r = 0; r < h; r++) bfly(x[R0 + r], x[R0 + r + h]);
#pragma unroll
        for (int k = 0; k < q; k++) x[R0 + h + q + k] = -x[R0 + h + q + k];
#pragma unroll
        for (int k = 0; k < q; k++) {
            if (k == 0) rotA<L, 0, 0>(x[R0 + h], x[R0 + h + q], twA, lane);
            else if (k == NR / 8 && NR >= 8) rotA<L, 0, (NR >= 8 ? NR / 8 : 0)>(x[R0 + h + k], x[R0 + h + q + k], twA, lane);
            else {
                const FftVec4 T0 = fftr_ldg4(twA + 4 * (FFTR_TWA_OFF(L, 0) + 32 * k + lane));
                rot(x[R0 + h + k], x[R0 + h + q + k], FftTwC{T0.x, T0.y, T0.z});
            }
        }
        rsA<L - 1, R0, N>(x, twA, lane);
        srA<L - 2, R0 + h, R0 + h + q, N>(x, twA, lane);
    }
}

// ---- phase 2 ---------------------------------------------------------------------------------------------------------
SIMT_FN void load_block(const float *X, int slot, float (&x)[32])
{
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const FftVec4 v = fftr_ld4(X + FFTR_SLOT_WORDS * slot + 4 * q);
        x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
    }
}
SIMT_FN void store_block(float *X, int slot, const float (&x)[32])
{
#pragma unroll
    for (int q = 0; q < 8; q++) fftr_st4(X + FFTR_SLOT_WORDS * slot + 4 * q, x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
}

// Both passes over the blocks the phase-1 code left in X[] (all lanes of the warp call this; X is the warp's own)
SIMT_FN void phase2(float *X, int lane)
{
    float x[32];
    lanes::sync();
    {   // pass 0: kind a, all 32 lanes
        load_block(X, lane, x);
        Half h;
        h.mask = 0xffffffffu; h.partner = c_fftr.partner[0][lane]; h.is_xi = c_fftr.is_xi[0][lane] != 0; h.sgn = h.is_xi ? 0x80000000u : 0u;
        sr_half<32, 0>(x, h);
        store_block(X, lane, x);
    }
    if (lane < 24) {   // pass 1: kind b in lanes 0..15, kind c in 16..19, kind d in 20..23
        load_block(X, 32 + lane, x);
        if (lane < 16) {
            Half h;
            h.mask = 0x0000ffffu; h.partner = c_fftr.partner[1][lane]; h.is_xi = c_fftr.is_xi[1][lane] != 0; h.sgn = h.is_xi ? 0x80000000u : 0u;
            sr_half_steps234<64, 0>(x, h);
            sr_half<16, 0>(x, h);
            sr_half<16, 16>(x, h);
        } else if (lane < 20) {
            rs_full<32, 0>(x);
        } else {
            rs_steps234<64, 0>(x);
            sr_full<16, 0, 16>(x);
        }
        store_block(X, 32 + lane, x);
    }
    lanes::sync();
}

}  // namespace fftr

// One granule-channel's transforms: in_long = 1024 windowed samples, in_short[t] = 256 windowed samples (phase-1 layout:
// the caller passes element lane + 32 r in xl[r] / xs[t][r]); results stay in X[] and are read through the output maps.
SIMT_FN void fft_regs_run(float (&xl)[32], float (&xs)[3][8], const float *twA, float *X, int lane)
{
    fftr::rsA<10, 0, 32>(xl, twA, lane);
#pragma unroll
    for (int r = 0; r < 32; r++) X[c_fftr.long_word[r] + lane] = xl[r];
#pragma unroll
    for (int t = 0; t < 3; t++) {
        fftr::rsA<8, 0, 8>(xs[t], twA, lane);
#pragma unroll
        for (int r = 0; r < 8; r++) X[c_fftr.short_word[t][r] + lane] = xs[t][r];
    }
    fftr::phase2(X, lane);
}

}  // namespace mp3gpu
