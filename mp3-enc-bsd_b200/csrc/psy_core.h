// psy_core.h — psychoacoustic model 2, Layer III branch (replaces L3psycho_anal(),
// /root/reference/src/l3psy.c:443-740, and fft()/enphinew(), subs.c:38-534).
//
// The reference call for granule g of a channel depends on earlier calls only through
//   (i)   r / phi of FFT lines 0..5 of calls g-1, g-2          (l3psy.c:498-501)
//   (ii)  nb of calls g-1, g-2 (pre-echo control)              (l3psy.c:632-635)
//   (iii) blocktype_old and the delayed ratio / ratio_s        (l3psy.c:452-456, 653-733)
// so the work is split in two kernels:
//   psy_front : one warp per granule-channel, fully parallel — windows, the 1024- and 3x256-point
//               FFTs (bit-exact FP32 op program, tables.h), energies, unpredictability of lines
//               6..205, partition energies, the energy spreading and the complete short-block
//               threshold path.  Writes a PsyMid record (1.5 KB).
//   psy_scan  : one warp per (stream, channel), sequential over granules — the few hundred flops
//               per granule that need history: cw of lines 0..5, tonality, SNR, nb, pre-echo,
//               PE, long-block ratios, block-type state machine.  Writes PsyOut.
// Mixed float/double accumulation types and orders follow the reference statement by statement.
#pragma once
#include "rate_loop_core.h"  // PsyOut
#include "simt.h"
#include "tables.h"
#include "fft_regs.h"

namespace mp3gpu {

using simt::PerThread;
using simt::WarpCtx;

struct FftDev {
    const uint32_t *words;  // op stream, see the class table in tables.h
    const int *seg_word;    // [4 * n_levels + 1] segment (level, class) -> offset into words[]
    int n_levels;
    const uint32_t *out;    // bin i (0..n/2) -> re | im << 16, each = slot of logical index i resp. n - i | (stored negated) << 15
};

struct PsyDev {
    const PsyTables *T;
    const FftTwiddle *tw;
    FftDev f1024, f256;
    // register FFT (fft_regs.h)
    const float *twA;
    const uint32_t *out_long, *out_short;   // [513], [3][132]
};

struct PsyMid {
    double eb[64];       // long partition energies (history free)
    float ecb[64];       // spread energy
    float cb[64];        // weighted unpredictability, valid for partitions >= n_hist_part
    float tail[48];      // energies of lines >= tail_l (they fold into partition 0), line order
    float e6[8], phi6[8];
    float es[3][132];    // energies of the three short transforms, lines 0..128 (the short-block ratios are formed from them
                         // by psy_scan, and only for granules that do switch to short blocks: the cold tail of the record)
};

// 6816 B per warp (8 warps per CTA, 4 CTAs per SM).  x[] holds the FFT working set: the 1024-point transform, then the three
// 256-point transforms side by side (data set sb at x + sb * FFT_BATCH_BYTES / 4: 264 skewed words + 128 dummy words).
// The short energies / phases stay IN PLACE (energy of bin i over re(i), phase over im(i) = re(256 - i); each pair is
// touched by one lane only) and are read through the output map; cwv and eb live in the (by then dead) dummy words of
// data sets 0 and 1; thr overlays E[] once the long partition energies have been formed (E is dead by then).
struct PsyFrontSmem {
    float x[FFT_X_ALLOC];
    float E[520];
    double prod[516];   // cw * e per line (psy_front_tail)
};
struct PsyFrontView {
    float *x, *E;
    double *cwv;        // [52]  at x + 264
    double *eb;         // [64]  at x + 656
    double *thr;        // [64]  at E + 0   (after the long partition energies)
    double *prod;
    SIMT_FN explicit PsyFrontView(PsyFrontSmem &S)
        : x(S.x), E(S.E), cwv(reinterpret_cast<double *>(S.x + 264)), eb(reinterpret_cast<double *>(S.x + 656)), thr(reinterpret_cast<double *>(S.E)), prod(S.prod) {}
};
static_assert(264 + 2 * 52 <= FFT_BATCH_BYTES / 4 && 656 >= FFT_BATCH_BYTES / 4 + 264 && 656 + 2 * 64 <= 2 * (FFT_BATCH_BYTES / 4), "overlays must sit in dummy words");

struct PsyScanSmem {
    double eb[64], thr[64], prod[64];
    double cw6[8];
    float cb[64];
    double pe;
};

struct PsyChanState {  // persistent per (stream, channel)
    float r1[8], r2[8], p1[8], p2[8];  // r / phi of lines 0..5 after calls g-1 (1) and g-2 (2)
    float nb1[64], nb2[64];
    double ratio_l[24];
    double ratio_s[36];
    int blocktype_old, pad;
};

static const double kLn2Log10 = 0.2302585093;  // LN_TO_LOG10, common.h:204

// ---- FFT op-program executor -------------------------------------------------------------------------
// A row is 32 ops of one class, one per lane; operands are byte offsets into the warp's x[].  The rows of a segment
// are independent (one dependency level), so a trip takes U rows, issues all their loads and only then their stores:
// the compiler cannot prove that the stores of one op do not alias the loads of the next, the level structure does.
// Padding ops of classes 0..2 work on per-lane dummy words (tables.h), so there is no NOP test in the fast loops.
SIMT_FN float fft_ld(const float *x, unsigned byte_off) { return *reinterpret_cast<const float *>(reinterpret_cast<const char *>(x) + byte_off); }
SIMT_FN void fft_st(float *x, unsigned byte_off, float v) { *reinterpret_cast<float *>(reinterpret_cast<char *>(x) + byte_off) = v; }
SIMT_FN float fft_flip(float v, unsigned sign_word)     // v with its sign flipped when bit 31 of sign_word is set
{
#if SIMT_DEV
    return __uint_as_float(__float_as_uint(v) ^ (sign_word & 0x80000000u));
#else
    return (sign_word & 0x80000000u) ? -v : v;
#endif
}

// NB transforms of the same length run through one pass over the op program: data set b lives FFT_BATCH_BYTES * b
// behind the first one, so a fetched and decoded op is applied NB times (the three short transforms of a granule).
template <int U, int NB>
SIMT_FN void fft_rows_bfly(const uint32_t *w, float *x)        // t=a+b; b=a-b; a=t
{
    float a[U][NB], b[U][NB];
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
        for (int n = 0; n < NB; n++) { a[u][n] = fft_ld(x, (w[u] & 0xffffu) + n * FFT_BATCH_BYTES); b[u][n] = fft_ld(x, (w[u] >> 16) + n * FFT_BATCH_BYTES); }
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
        for (int n = 0; n < NB; n++) {
            fft_st(x, (w[u] & 0xffffu) + n * FFT_BATCH_BYTES, simt::fadd(a[u][n], b[u][n]));
            fft_st(x, (w[u] >> 16) + n * FFT_BATCH_BYTES, simt::fsub(a[u][n], b[u][n]));
        }
}

struct FftW2 { uint32_t lo, hi; };

template <int U, int NB>
SIMT_FN void fft_rows_cross(const FftW2 *w, float *x)          // t1=a+d; t2=c+b; c=c-b; b=a-d; a=t1; d=t2
{
    float a[U][NB], b[U][NB], c[U][NB], d[U][NB];
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
        for (int n = 0; n < NB; n++) {
            a[u][n] = fft_ld(x, (w[u].lo & 0xffffu) + n * FFT_BATCH_BYTES); b[u][n] = fft_ld(x, (w[u].lo >> 16) + n * FFT_BATCH_BYTES);
            c[u][n] = fft_ld(x, (w[u].hi & 0xffffu) + n * FFT_BATCH_BYTES); d[u][n] = fft_ld(x, (w[u].hi >> 16) + n * FFT_BATCH_BYTES);
        }
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
        for (int n = 0; n < NB; n++) {
            fft_st(x, (w[u].lo & 0xffffu) + n * FFT_BATCH_BYTES, simt::fadd(a[u][n], d[u][n]));
            fft_st(x, (w[u].lo >> 16) + n * FFT_BATCH_BYTES, simt::fsub(a[u][n], d[u][n]));
            fft_st(x, (w[u].hi & 0xffffu) + n * FFT_BATCH_BYTES, simt::fsub(c[u][n], b[u][n]));
            fft_st(x, (w[u].hi >> 16) + n * FFT_BATCH_BYTES, simt::fadd(c[u][n], b[u][n]));
        }
}

struct alignas(16) FftW4 { uint32_t lo; float cn, spcn, smcn; };

template <int U, int NB>
SIMT_FN void fft_rows_rot(const FftW4 *w, float *x)            // t2=cn*(a+c); t1=spcn*a+t2; a=smcn*c+t2; c=t1
{
    float a[U][NB], c[U][NB];
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
        for (int n = 0; n < NB; n++) {
            a[u][n] = fft_ld(x, (w[u].lo & 0xffffu) + n * FFT_BATCH_BYTES);
            c[u][n] = fft_flip(fft_ld(x, ((w[u].lo >> 16) & 0x7fffu) + n * FFT_BATCH_BYTES), w[u].lo);
        }
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
        for (int n = 0; n < NB; n++) {
            const float t2 = simt::fmul(w[u].cn, simt::fadd(a[u][n], c[u][n]));
            fft_st(x, (w[u].lo & 0xffffu) + n * FFT_BATCH_BYTES, simt::fadd(simt::fmul(w[u].smcn, c[u][n]), t2));
            fft_st(x, ((w[u].lo >> 16) & 0x7fffu) + n * FFT_BATCH_BYTES, simt::fadd(simt::fmul(w[u].spcn, a[u][n]), t2));
        }
}

template <int NB>
SIMT_FN void fft_row_misc(const FftW2 w, float *x)             // the rare shapes, one row at a time
{
    const double SQ = 0.707106781186547524401;  // SQHALF, subs.c:26
    const int type = (int)(w.hi & 7u);
    if (type == FFT_NOP) return;
#pragma unroll
    for (int n = 0; n < NB; n++) {
        const unsigned oa = (w.lo & 0xffffu) + n * FFT_BATCH_BYTES, oc = (w.lo >> 16) + n * FFT_BATCH_BYTES;
        const float a = fft_flip(fft_ld(x, oa), w.hi << 23), c = fft_flip(fft_ld(x, oc), w.hi << 22);
        float ra, rc;
        if (type == FFT_BFLY) { ra = simt::fadd(a, c); rc = simt::fsub(a, c); }
        else if (type == FFT_ROT8A) {                                 // t1=SQ*(a+c); c=SQ*(c-a); a=t1 (double multiply)
            ra = (float)simt::dmul(SQ, (double)simt::fadd(a, c)); rc = (float)simt::dmul(SQ, (double)simt::fsub(c, a));
        } else {                                                      // FFT_ROT8B: t2=SQ*(c-a); c=-SQ*(a+c); a=t2
            ra = (float)simt::dmul(SQ, (double)simt::fsub(c, a)); rc = (float)simt::dmul(-SQ, (double)simt::fadd(a, c));
        }
        fft_st(x, oa, ra);
        fft_st(x, oc, rc);
    }
}

// U rows per trip, NB data sets; dummy_word: first of the 128 padding-lane words of data set 0 (n + n / 32)
template <int U, int NB>
SIMT_FN void fft_run(const WarpCtx &w, const FftDev &P, const FftTwiddle *tw, float *x, int dummy_word)
{
    FOR_THREADS(w)
    for (int n = 0; n < NB; n++)
        for (int j = 0; j < 4; j++) x[n * (FFT_BATCH_BYTES / 4) + dummy_word + 32 * j + lane] = 0.0f;   // the padding lanes' dummy words
    END_THREADS
    w.sync();
    for (int l = 0; l < P.n_levels; l++) {
        const int *sg = P.seg_word + 4 * l;
        const int s0 = sg[0], s1 = sg[1], s2 = sg[2], s3 = sg[3], s4 = sg[4];
        FOR_THREADS(w)
        {
            // Op words are fetched one trip ahead (the first trip of all four class loops up front): a trip otherwise
            // starts with a global-load round trip that nothing in the warp can overlap (ncu: 25 % of the stall samples).
            // Fetches run past the end of a segment on purpose (into the next segment / the pad words behind the program).
            // 32-bit word indices: a 64-bit pointer loop costs ~6 more integer instructions per trip.
            const uint32_t *W = P.words;
            const FftW2 *W2 = reinterpret_cast<const FftW2 *>(P.words);
            const FftW4 *W4 = reinterpret_cast<const FftW4 *>(P.words);
            int ib = s0 + lane, ic = (s1 >> 1) + lane, ir = (s2 >> 2) + lane, im = (s3 >> 1) + lane;
            uint32_t wb[U] = {}; FftW2 wc[U] = {}, wm = {}; FftW4 wr[U] = {};
            // (empty segments are not fetched: the look-ahead words cost L2 bandwidth, which this kernel is short of)
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (s1 > s0) wb[u] = W[ib + 32 * u];
                if (s2 > s1) wc[u] = W2[ic + 32 * u];
                if (s3 > s2) wr[u] = W4[ir + 32 * u];
            }
            if (s4 > s3) wm = W2[im];
            for (; ib + 32 * (U - 1) < s1; ib += 32 * U) {
                uint32_t nx[U];
#pragma unroll
                for (int u = 0; u < U; u++) nx[u] = W[ib + 32 * (U + u)];
                fft_rows_bfly<U, NB>(wb, x);
#pragma unroll
                for (int u = 0; u < U; u++) wb[u] = nx[u];
            }
#pragma unroll
            for (int u = 0; u < U - 1; u++) if (ib + 32 * u < s1) fft_rows_bfly<1, NB>(wb + u, x);     // the up to U - 1 rows left
            for (; ic + 32 * (U - 1) < (s2 >> 1); ic += 32 * U) {
                FftW2 nx[U];
#pragma unroll
                for (int u = 0; u < U; u++) nx[u] = W2[ic + 32 * (U + u)];
                fft_rows_cross<U, NB>(wc, x);
#pragma unroll
                for (int u = 0; u < U; u++) wc[u] = nx[u];
            }
#pragma unroll
            for (int u = 0; u < U - 1; u++) if (ic + 32 * u < (s2 >> 1)) fft_rows_cross<1, NB>(wc + u, x);
            for (; ir + 32 * (U - 1) < (s3 >> 2); ir += 32 * U) {
                FftW4 nx[U];
#pragma unroll
                for (int u = 0; u < U; u++) nx[u] = W4[ir + 32 * (U + u)];
                fft_rows_rot<U, NB>(wr, x);
#pragma unroll
                for (int u = 0; u < U; u++) wr[u] = nx[u];
            }
#pragma unroll
            for (int u = 0; u < U - 1; u++) if (ir + 32 * u < (s3 >> 2)) fft_rows_rot<1, NB>(wr + u, x);
            for (; im < (s4 >> 1); im += 32) {
                const FftW2 nx = W2[im + 32];
                fft_row_misc<NB>(wm, x);
                wm = nx;
            }
        }
        END_THREADS
        w.sync();
    }
}

// energy / phase of bin i of short transform sb, left in place by psy_front (energy over re, phase over im)
SIMT_FN float short_energy(const FftDev &P, const float *x, int sb, int i) { return x[sb * (FFT_BATCH_BYTES / 4) + (P.out[i] & 0x7fffu)]; }
SIMT_FN float short_phase(const FftDev &P, const float *x, int sb, int i) { return x[sb * (FFT_BATCH_BYTES / 4) + ((P.out[i] >> 16) & 0x7fffu)]; }

// energy + phase of bin i (enphinew, subs.c:53-123); n = transform length; om = P.out[i]
SIMT_FN void bin_energy_phase(unsigned om, const float *x, int n, int i, bool want_phi, float *e_out, float *phi_out)
{
    float re = x[om & 0x7fffu];
    if (om & 0x8000u) re = -re;
    if (i == 0 || i == n / 2) {
        *e_out = simt::fmul(re, re);
        if (want_phi) *phi_out = (float)atan2(0.0, (double)re);
        return;
    }
    float im = x[(om >> 16) & 0x7fffu];
    if (om & 0x80000000u) im = -im;
    float e = simt::fadd(simt::fmul(re, re), simt::fmul(im, im));
    if ((double)e < 0.0005) { *e_out = (float)0.0005; if (want_phi) *phi_out = 0.0f; }
    else { *e_out = e; if (want_phi) *phi_out = (float)atan2(-(double)im, (double)re); }
}

SIMT_FN double unpredictability(double r_new, double phi_new, double r_prime, double phi_prime)
{
    // l3psy.c:503-511 / 540-547
    double t1 = simt::dsub(simt::dmul(r_new, cos(phi_new)), simt::dmul(r_prime, cos(phi_prime)));
    double t2 = simt::dsub(simt::dmul(r_new, sin(phi_new)), simt::dmul(r_prime, sin(phi_prime)));
    double t3 = simt::dadd(r_new, fabs(r_prime));
    if (t3 != 0.0) return sqrt(simt::dadd(simt::dmul(t1, t1), simt::dmul(t2, t2))) / t3;
    return 0.0;
}

// ---------------------------------------------------------------------------------------------------
// history-free tail of psy_front: partition energies, weighted unpredictability, energy spreading.
// E[0..512]: energies of the long transform in line order; cwv[i]: unpredictability of lines 6 + 4 i .. 9 + 4 i (i < 50);
// eb_s[64], prod[513]: scratch (all in the warp's shared memory)
// ---------------------------------------------------------------------------------------------------
SIMT_FN void psy_front_tail(const WarpCtx &w, const PsyTables &T, const float *E, const double *cwv, double *eb_s, double *prod, PsyMid *out)
{
    FOR_THREADS(w)
    for (int j = T.tail_l + lane; j <= 512; j += 32) out->tail[j - T.tail_l] = E[j];
    // cw[j] * e[j] of every line, l3psy.c:574-577 (lines 0..5 belong to the partitions whose cb psy_scan forms from history;
    // lines >= 206 have cw = 0.4): the partition chains below are then three conversions and two additions per line
#pragma unroll 1
    for (int j = lane; j <= 512; j += 32) {
        const double cw = (j < 6) ? 0.0 : (j < 206) ? cwv[(j - 6) >> 2] : 0.4;
        prod[j] = simt::dmul(cw, (double)E[j]);
    }
    END_THREADS
    w.sync();
    // partition energy / weighted unpredictability, l3psy.c:565-578: eb is a double, cb a FLOAT accumulator, so both are
    // sequential chains in line order; a lane runs two chains at a time (slot 0: partition lane, slot 1: partition lane + 32).
    // Partition 0 also takes the lines >= tail_l (zero-initialised partition map, l3psy.c:93): that chain runs in the idle
    // slot 1 of lane 31, started from the sum of partition 0's own lines, beside the wide partitions instead of after them.
    FOR_THREADS(w)
    {
        int j0 = T.ch_lo[0][lane], j1 = T.ch_lo[1][lane];
        const int h0 = T.ch_hi[0][lane], h1 = T.ch_hi[1][lane];
        double eb0 = 0.0, eb1 = 0.0;
        float cb0 = 0.0f, cb1 = 0.0f;
        if (lane == 31) for (int j = T.lo_l[0]; j < T.hi_l[0]; j++) eb1 = simt::dadd(eb1, (double)E[j]);
        const int wa = T.ch_wmax[0], wb = T.ch_wmax[1];
#pragma unroll 1
        for (int i = 0; i < wa; i++) {
            if (j0 < h0) { eb0 = simt::dadd(eb0, (double)E[j0]); cb0 = (float)simt::dadd((double)cb0, prod[j0]); j0++; }
            if (j1 < h1) { eb1 = simt::dadd(eb1, (double)E[j1]); cb1 = (float)simt::dadd((double)cb1, prod[j1]); j1++; }
        }
#pragma unroll 4
        for (int i = wa; i < wb; i++)
            if (j1 < h1) { eb1 = simt::dadd(eb1, (double)E[j1]); cb1 = (float)simt::dadd((double)cb1, prod[j1]); j1++; }
        if (lane > 0) { eb_s[lane] = eb0; out->eb[lane] = eb0; }
        const int p1 = (lane == 31) ? 0 : lane + 32;
        eb_s[p1] = eb1; out->eb[p1] = eb1;
        out->cb[lane] = cb0;
        out->cb[lane + 32] = (lane == 31) ? 0.0f : cb1;
        if (lane == 31) { eb_s[63] = 0.0; out->eb[63] = 0.0; }
    }
    END_THREADS
    w.sync();
    // energy spreading, l3psy.c:586-605 (ecb is a float accumulator)
    FOR_THREADS(w)
    {
        // The two partitions of a lane (b and b + 32) advance together: two independent float <- double accumulation
        // chains in flight.
        const int b0 = lane, b1 = lane + 32;
        const int lo0 = T.spr_lo[b0], n0 = T.spr_hi[b0] - lo0 + 1;
        const int lo1 = (b1 < 63) ? T.spr_lo[b1] : 0, n1 = (b1 < 63) ? T.spr_hi[b1] - lo1 + 1 : 0;
        const bool sparse = T.sparse != 0;
        const double *s0p = T.s3_band + b0, *s1p = T.s3_band + b1, *e0p = eb_s + lo0, *e1p = eb_s + lo1;
        float e0 = 0.0f, e1 = 0.0f;
        const int wmax = T.spr_wmax;
#pragma unroll 4
        for (int i = 0; i < wmax; i++) {
            // step i of each lane's own row range [lo, hi] (banded matrix layout: one coalesced request per step); the terms
            // are added in the reference's order, k ascending
            if (i < n0) { const double s = s0p[i * 64]; if (sparse || s != 1.0) e0 = (float)simt::dadd((double)e0, simt::dmul(s, e0p[i])); }
            if (i < n1) { const double s = s1p[i * 64]; if (sparse || s != 1.0) e1 = (float)simt::dadd((double)e1, simt::dmul(s, e1p[i])); }
        }
        out->ecb[b0] = e0;
        out->ecb[b1] = e1;
    }
    END_THREADS
    w.sync();
}

// ---------------------------------------------------------------------------------------------------
// psy_front: history-free part, one warp per granule-channel.
// pcm points at the first NEW sample of the granule (sample 576 g); pcm[-768 ..  575] are read.
// ---------------------------------------------------------------------------------------------------
SIMT_FN void psy_front(const WarpCtx &w, const PsyDev &D, PsyFrontSmem &S, const short *pcm, PsyMid *out)
{
    const PsyTables &T = *D.T;
    PsyFrontView M(S);
    // long window + FFT, l3psy.c:483-494
    FOR_THREADS(w)
#pragma unroll 8
    for (int j = lane; j < 1024; j += 32) M.x[FFT_SKEW(j)] = simt::fmul(T.hann_l[j], (float)(int)pcm[j - 768]);
    END_THREADS
    w.sync();
#ifndef FFT_LONG_U
#define FFT_LONG_U 2     // rows per trip of the 1024-point transform
#endif
    fft_run<FFT_LONG_U, 1>(w, D.f1024, D.tw, M.x, FFT_X_WORDS);
    FOR_THREADS(w)
#pragma unroll 4
    for (int i = lane; i <= 512; i += 32) {           // energies of all bins: four independent map-load -> x-load chains in flight
        float e, ph;
        bin_energy_phase(D.f1024.out[i], M.x, 1024, i, false, &e, &ph);
        M.E[i] = e;
    }
    if (lane < 6) {                                   // lines 0..5 also need the phase (psy_scan predicts them from history)
        float e, ph = 0.f;
        bin_energy_phase(D.f1024.out[lane], M.x, 1024, lane, true, &e, &ph);
        out->e6[lane] = e; out->phi6[lane] = ph;
    }
    END_THREADS
    w.sync();
    // three short FFTs, l3psy.c:518-527: one pass over the 256-point program transforms all three windows
    {
        FOR_THREADS(w)
        for (int sb = 0; sb < 3; sb++)
#pragma unroll
            for (int j = lane; j < 256; j += 32)
                M.x[sb * (FFT_BATCH_BYTES / 4) + FFT_SKEW(j)] = simt::fmul(T.hann_s[j], (float)(int)pcm[j - 768 + 128 * (2 + sb)]);
        END_THREADS
        w.sync();
        fft_run<1, 3>(w, D.f256, D.tw, M.x, 256 + 8);
        FOR_THREADS(w)
        for (int sb = 0; sb < 3; sb++) {
            float *xs = M.x + sb * (FFT_BATCH_BYTES / 4);
            for (int i = lane; i <= 128; i += 32) {
                float e, ph = 0.f;
                const bool wp = (i >= 2 && i < 52);
                const unsigned om = D.f256.out[i];
                bin_energy_phase(om, xs, 256, i, wp, &e, &ph);
                xs[om & 0x7fffu] = e;
                out->es[sb][i] = e;
                if (wp) xs[(om >> 16) & 0x7fffu] = ph;
            }
        }
        END_THREADS
        w.sync();
    }
    // unpredictability of lines 6..205 in groups of four, l3psy.c:531-549
    FOR_THREADS(w)
    for (int i = lane; i < 50; i += 32) {
        const int k = i + 2;
        double r_prime = simt::dsub(simt::dmul(2.0, sqrt((double)short_energy(D.f256, M.x, 0, k))), sqrt((double)short_energy(D.f256, M.x, 2, k)));
        double phi_prime = simt::dsub(simt::dmul(2.0, (double)short_phase(D.f256, M.x, 0, k)), (double)short_phase(D.f256, M.x, 2, k));
        double r2 = sqrt((double)short_energy(D.f256, M.x, 1, k));
        M.cwv[i] = unpredictability(r2, (double)short_phase(D.f256, M.x, 1, k), r_prime, phi_prime);
    }
    END_THREADS
    w.sync();
    psy_front_tail(w, T, M.E, M.cwv, M.eb, M.prod, out);
}

#if SIMT_DEV
// ---------------------------------------------------------------------------------------------------
// psy_front with the transforms in registers (fft_regs.h): same results as psy_front, bit for bit.
// X: the warp's FFTR_X_WORDS floats of shared memory.  Once the energies are formed the transform data is dead and X is
// reused: E[0..512] at X, cwv[52] (double) at X + 520, eb[64] (double) at X + 624.
// ---------------------------------------------------------------------------------------------------
#define PSYF2_CWV_WORD 520
#define PSYF2_EB_WORD 624
#define PSYF2_PROD_WORD 752
static_assert(PSYF2_PROD_WORD + 2 * 516 <= FFTR_X_WORDS && PSYF2_EB_WORD + 128 <= PSYF2_PROD_WORD && PSYF2_CWV_WORD + 104 <= PSYF2_EB_WORD, "overlays");
SIMT_FN void psy_front_regs(const WarpCtx &w, const PsyDev &D, float *X, const short *pcm, PsyMid *out)
{
    const PsyTables &T = *D.T;
    const int lane = w.lane;
    {
        float xl[32], xs[3][8];
        // long window, l3psy.c:483-494; three short windows, l3psy.c:518-527
#pragma unroll
        for (int r = 0; r < 32; r++) xl[r] = simt::fmul(T.hann_l[lane + 32 * r], (float)(int)pcm[lane + 32 * r - 768]);
#pragma unroll
        for (int t = 0; t < 3; t++)
#pragma unroll
            for (int r = 0; r < 8; r++) xs[t][r] = simt::fmul(T.hann_s[lane + 32 * r], (float)(int)pcm[lane + 32 * r - 768 + 128 * (2 + t)]);
        fft_regs_run(xl, xs, D.twA, X, lane);
    }
    // energies of the long transform: bins lane + 32 k (k < 16) and bin 512; phases of lines 0..5
    float e[17];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        float ph;
        bin_energy_phase(D.out_long[lane + 32 * k], X, 1024, lane + 32 * k, false, &e[k], &ph);
    }
    e[16] = 0.f;
    if (lane == 0) { float ph; bin_energy_phase(D.out_long[512], X, 1024, 512, false, &e[16], &ph); }
    // Phases (double-precision atan2, the most expensive single operation of the kernel): lines 2..51 of the three short
    // transforms and lines 0..5 of the long one = 156 bins, spread over five passes of the 32 lanes (one copy of the code).
    // The short phases go to the pad words of the transform slots (item q -> pad word q % 4 of slot q / 4).
#pragma unroll 1
    for (int c = 0; c < 5; c++) {
        const int q = lane + 32 * c;
        if (q < 156) {
            const bool is_long = q >= 150;
            const int i3 = (q * 171) >> 9, t = q - 3 * i3;                 // q / 3, q % 3 (q < 256)
            const int k = is_long ? q - 150 : i3 + 2;
            float en, ph = 0.f;
            bin_energy_phase(is_long ? D.out_long[k] : D.out_short[132 * t + k], X, is_long ? 1024 : 256, k, true, &en, &ph);
            if (is_long) { out->e6[k] = en; out->phi6[k] = ph; }
            else { out->es[t][k] = en; X[FFTR_SLOT_WORDS * (q >> 2) + 32 + (q & 3)] = ph; }
        }
    }
    w.sync();
    // unpredictability of lines 6..205 in groups of four, l3psy.c:531-549
    double cw[2] = {0.0, 0.0};
#pragma unroll
    for (int it = 0; it < 2; it++) {
        const int i = lane + 32 * it;
        if (i < 50) {
            const int k = i + 2;
            float es[3], ps[3];
#pragma unroll
            for (int t = 0; t < 3; t++) {
                float ph;
                bin_energy_phase(D.out_short[132 * t + k], X, 256, k, false, &es[t], &ph);      // the energy again: 3 flops
                const int q = 3 * i + t;
                ps[t] = X[FFTR_SLOT_WORDS * (q >> 2) + 32 + (q & 3)];
            }
            const double r_prime = simt::dsub(simt::dmul(2.0, sqrt((double)es[0])), sqrt((double)es[2]));
            const double phi_prime = simt::dsub(simt::dmul(2.0, (double)ps[0]), (double)ps[2]);
            cw[it] = unpredictability(sqrt((double)es[1]), (double)ps[1], r_prime, phi_prime);
        }
    }
    // the other short lines (0, 1, 52..128): energies only
#pragma unroll
    for (int it = 0; it < 3; it++) {
        const int idx = lane + 32 * it, k = idx < 2 ? idx : idx + 50;
        if (k <= 128) {
#pragma unroll
            for (int t = 0; t < 3; t++) {
                float es, ph;
                bin_energy_phase(D.out_short[132 * t + k], X, 256, k, false, &es, &ph);
                out->es[t][k] = es;
            }
        }
    }
    w.sync();                                           // every read of the transform data is done
    double *cwv = reinterpret_cast<double *>(X + PSYF2_CWV_WORD), *eb = reinterpret_cast<double *>(X + PSYF2_EB_WORD);
#pragma unroll
    for (int k = 0; k < 16; k++) X[lane + 32 * k] = e[k];
    if (lane == 0) X[512] = e[16];
    cwv[lane] = cw[0];
    if (lane < 18) cwv[lane + 32] = cw[1];
    w.sync();
    psy_front_tail(w, T, X, cwv, eb, reinterpret_cast<double *>(X + PSYF2_PROD_WORD), out);
}
#endif

// ---------------------------------------------------------------------------------------------------
// psy_scan: history-dependent part, one warp per (stream, channel), granules in order.
// ---------------------------------------------------------------------------------------------------
struct PsyScanRegs {
    PerThread<float> r1, r2, p1, p2;  // lane = FFT line (< 6)
    PerThread<float> nb1[2], nb2[2];  // partition lane + 32 h
    PerThread<double> rl;             // long ratio, lane = sfb (< 21)
    PerThread<double> rs[2];          // short ratio, index lane + 32 h (< 36)
    int blocktype_old;
};

SIMT_FN void psy_scan_load(const WarpCtx &w, const PsyChanState &S, PsyScanRegs &R)
{
    FOR_THREADS(w)
    R.r1() = (lane < 6) ? S.r1[lane] : 0.f; R.r2() = (lane < 6) ? S.r2[lane] : 0.f;
    R.p1() = (lane < 6) ? S.p1[lane] : 0.f; R.p2() = (lane < 6) ? S.p2[lane] : 0.f;
#pragma unroll
    for (int h = 0; h < 2; h++) { R.nb1[h]() = S.nb1[lane + 32 * h]; R.nb2[h]() = S.nb2[lane + 32 * h]; }
    R.rl() = (lane < 21) ? S.ratio_l[lane] : 0.0;
    R.rs[0]() = S.ratio_s[lane];
    R.rs[1]() = (lane < 4) ? S.ratio_s[32 + lane] : 0.0;
    END_THREADS
    R.blocktype_old = S.blocktype_old;
}

SIMT_FN void psy_scan_store(const WarpCtx &w, PsyChanState &S, const PsyScanRegs &R)
{
    FOR_THREADS(w)
    if (lane < 6) { S.r1[lane] = R.r1(); S.r2[lane] = R.r2(); S.p1[lane] = R.p1(); S.p2[lane] = R.p2(); }
#pragma unroll
    for (int h = 0; h < 2; h++) { S.nb1[lane + 32 * h] = R.nb1[h](); S.nb2[lane + 32 * h] = R.nb2[h](); }
    if (lane < 21) S.ratio_l[lane] = R.rl();
    S.ratio_s[lane] = R.rs[0]();
    if (lane < 4) S.ratio_s[32 + lane] = R.rs[1]();
    if (lane == 0) S.blocktype_old = R.blocktype_old;
    END_THREADS
}

// Short-block ratios of one granule, l3psy.c:698-729 (uses the LONG spreading matrix and norm_l, sic), from the short
// energies psy_front left in PsyMid.  Only granules that switch to short blocks need them (the reference computes them in
// the pe >= 1800 branch only), so psy_scan calls this on demand (out of line: the scan's common path stays small).
// eb / thr: 64 doubles of scratch each; ratio: [36] out.
SIMT_NOINLINE void psy_short_ratios(const WarpCtx &w, const PsyTables &T, const PsyMid &mid, double *eb_s, double *thr_s, double *ratio)
{
    for (int sb = 0; sb < 3; sb++) {
        const float *es = mid.es[sb];
        FOR_THREADS(w)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int p = lane + 32 * h;
            if (p < 42) {
                double eb = 0.0;
                if (p < T.n_s) for (int j = T.lo_s[p]; j < T.hi_s[p]; j++) eb = simt::dadd(eb, (double)es[j]);
                if (p == 0) for (int j = T.tail_s; j <= 128; j++) eb = simt::dadd(eb, (double)es[j]);
                eb_s[p] = eb;
            }
        }
        END_THREADS
        w.sync();
        FOR_THREADS(w)
        {
            const int b0 = lane, b1 = lane + 32;            // partitions 0..41: b1 is live in lanes 0..9 only
            float e0 = 0.0f, e1 = 0.0f;
            for (int k = 0; k < 42; k++) {
                const double ebk = eb_s[k];
                e0 = (float)simt::dadd((double)e0, simt::dmul(T.s3_lT[k * 64 + b0], ebk));
                e1 = (float)simt::dadd((double)e1, simt::dmul(T.s3_lT[k * 64 + b1], ebk));
            }
            const float nb0 = (float)simt::dmul(simt::dmul((double)e0, T.norm_l[b0]), T.snr_s_exp[b0]);
            thr_s[b0] = (T.qthr_s[b0] > (double)nb0) ? T.qthr_s[b0] : (double)nb0;
            if (b1 < 42) {
                const float nb1 = (float)simt::dmul(simt::dmul((double)e1, T.norm_l[b1]), T.snr_s_exp[b1]);
                thr_s[b1] = (T.qthr_s[b1] > (double)nb1) ? T.qthr_s[b1] : (double)nb1;
            }
        }
        END_THREADS
        w.sync();
        FOR_THREADS(w)
        if (lane < 12) {
            const int bu = T.bu_s[lane], bo = T.bo_s[lane];
            double en = simt::dadd(simt::dmul(T.w1_s[lane], eb_s[bu]), simt::dmul(T.w2_s[lane], eb_s[bo]));
            double thm = simt::dadd(simt::dmul(T.w1_s[lane], thr_s[bu]), simt::dmul(T.w2_s[lane], thr_s[bo]));
            for (int b = bu + 1; b < bo; b++) { en = simt::dadd(en, eb_s[b]); thm = simt::dadd(thm, thr_s[b]); }
            ratio[lane * 3 + sb] = (en != 0.0) ? thm / en : 0.0;
        }
        END_THREADS
        w.sync();
    }
}

SIMT_FN void psy_scan_step(const WarpCtx &w, const PsyTables &T, PsyScanSmem &M, const PsyMid &mid, PsyScanRegs &R, PsyOut *out)
{
    // delayed outputs first: the caller gets the ratios left by the PREVIOUS call (l3psy.c:452-456)
    FOR_THREADS(w)
    if (lane < 21) out->ratio_l[lane] = R.rl();
    out->ratio_s[lane] = R.rs[0]();
    if (lane < 4) out->ratio_s[32 + lane] = R.rs[1]();
    END_THREADS
    // unpredictability of lines 0..5, l3psy.c:496-512
    FOR_THREADS(w)
    if (lane < 6) {
        double r_prime = simt::dsub(simt::dmul(2.0, (double)R.r1()), (double)R.r2());
        double phi_prime = simt::dsub(simt::dmul(2.0, (double)R.p1()), (double)R.p2());
        float rn = (float)sqrt((double)mid.e6[lane]);
        float pn = mid.phi6[lane];
        M.cw6[lane] = unpredictability((double)rn, (double)pn, r_prime, phi_prime);
        R.r2() = R.r1(); R.r1() = rn;
        R.p2() = R.p1(); R.p1() = pn;
    }
    // 0.4 * e of the lines that fold into partition 0 (l3psy.c:574-577 with cw = 0.4), formed by all lanes into M.prod: the
    // sequential float accumulation below then reads shared memory instead of one dependent global load per line
    for (int j = T.tail_l + lane; j <= 512; j += 32) M.prod[j - T.tail_l] = simt::dmul(0.4, (double)mid.tail[j - T.tail_l]);
    END_THREADS
    w.sync();
    // cb of the partitions that hold lines 0..5 (+ the partition-0 tail), l3psy.c:570-578
    FOR_THREADS(w)
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int p = lane + 32 * h;
        float cb = mid.cb[p];
        if (p < T.n_hist_part) {
            cb = 0.0f;
            for (int j = T.lo_l[p]; j < T.hi_l[p]; j++) cb = (float)simt::dadd((double)cb, simt::dmul(M.cw6[j], (double)mid.e6[j]));
            if (p == 0)
                for (int j = 0; j <= 512 - T.tail_l; j++) cb = (float)simt::dadd((double)cb, M.prod[j]);
        }
        M.cb[p] = cb;
    }
    END_THREADS
    w.sync();
    // spreading of cb, tonality, SNR, nb, pre-echo, PE terms: l3psy.c:586-645
    FOR_THREADS(w)
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int b = lane + 32 * h;
        if (b < 63) {
            double ctb = 0.0;
            const int klo = T.spr_lo[b];
#pragma unroll 4
            for (int k = klo; k <= T.spr_hi[b]; k++) {   // latency-bound scan over the lane's own row range (banded layout)
                const double s = T.s3_band[(k - klo) * 64 + b];
                if (T.sparse || s != 1.0) ctb = simt::dadd(ctb, simt::dmul(s, (double)M.cb[k]));
            }
            const float ecb = mid.ecb[b];
            double cbb;
            if ((double)ecb != 0.0) {
                cbb = ctb / (double)ecb;
                if (cbb < 0.01) cbb = 0.01;
                cbb = log(cbb);
            } else cbb = 0.0;
            double tbb = simt::dsub(-0.299, simt::dmul(0.43, cbb));
            tbb = (0.0 > tbb) ? 0.0 : tbb;
            tbb = (1.0 < tbb) ? 1.0 : tbb;
            double v = simt::dadd(simt::dmul(29.0, tbb), simt::dmul(6.0, simt::dsub(1.0, tbb)));
            double snr = (T.minval[b] > v) ? T.minval[b] : v;
            float nb = (float)simt::dmul(simt::dmul((double)ecb, T.norm_l[b]), exp(simt::dmul(-snr, kLn2Log10)));
            double a = simt::dmul(2.0, (double)R.nb1[h]()), c = simt::dmul(16.0, (double)R.nb2[h]());
            double mn = (a < c) ? a : c;
            double t = ((double)nb < mn) ? (double)nb : mn;
            double thr = (T.qthr_l[b] > t) ? T.qthr_l[b] : t;
            R.nb2[h]() = R.nb1[h]();
            R.nb1[h]() = nb;
            const double eb = mid.eb[b];
            double l = log(simt::dadd(thr, 1.0) / simt::dadd(eb, 1.0));
            double tp = (0.0 < l) ? 0.0 : l;
            M.prod[b] = simt::dmul((double)T.numlines_pe[b], tp);
            M.thr[b] = thr;
            M.eb[b] = eb;
        }
    }
    END_THREADS
    w.sync();
    FOR_THREADS(w)
    if (lane == 0) {
        double pe = 0.0;
        for (int b = 0; b < 63; b++) pe = simt::dsub(pe, M.prod[b]);
        M.pe = pe;
    }
    END_THREADS
    w.sync();
    const double pe = M.pe;
    int blocktype;
    if (pe < 1800) {  // l3psy.c:651-685
        blocktype = (R.blocktype_old == 2) ? 3 : 0;
        FOR_THREADS(w)
        if (lane < 21) {
            const int bu = T.bu_l[lane], bo = T.bo_l[lane];
            double en = simt::dadd(simt::dmul(T.w1_l[lane], M.eb[bu]), simt::dmul(T.w2_l[lane], M.eb[bo]));
            double thm = simt::dadd(simt::dmul(T.w1_l[lane], M.thr[bu]), simt::dmul(T.w2_l[lane], M.thr[bo]));
            for (int b = bu + 1; b < bo; b++) { en = simt::dadd(en, M.eb[b]); thm = simt::dadd(thm, M.thr[b]); }
            R.rl() = (en != 0.0) ? thm / en : 0.0;
        }
        END_THREADS
    } else {          // l3psy.c:686-730
        blocktype = 2;
        if (R.blocktype_old == 0) R.blocktype_old = 1;
        if (R.blocktype_old == 3) R.blocktype_old = 2;
        psy_short_ratios(w, T, mid, M.eb, M.thr, M.prod);
        FOR_THREADS(w)
        R.rs[0]() = M.prod[lane];
        if (lane < 4) R.rs[1]() = M.prod[32 + lane];
        END_THREADS
    }
    FOR_THREADS(w)
    if (lane == 0) { out->pe = pe; out->block_type = R.blocktype_old; out->pad = 0; }
    END_THREADS
    R.blocktype_old = blocktype;  // l3psy.c:732-733
    w.sync();
}

}  // namespace mp3gpu
