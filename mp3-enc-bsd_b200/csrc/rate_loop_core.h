// rate_loop_core.h — Layer III rate loop for ONE stream, executed by ONE warp.
//
// Replaces iteration_loop()/outer_loop()/inner_loop()/bin_search_StepSize()/quantize()/count_bits()
// and helpers of /root/reference/src/loop.c (+ the reservoir recurrence of reservoir.c) for the batched
// C-ABI entry point mp3gpu_iteration_loop_batch().  One warp owns one stream and walks its
// (frame, granule, channel) sequence in order, because max_bits of a granule depends on the
// reservoir left by the previous one (reservoir.c:101-145).  Inside a granule all 576 lines are
// processed in parallel: each lane keeps 9 coefficient PAIRS in registers, and every bit count,
// maximum and table choice is a warp reduction.
//
// Layouts (k = 0..8, lane = 0..31, slot s = lane + 32 k in [0,288)):
//   long / start / stop blocks: slot s holds elements (2s, 2s+1)            -> Huffman pair s
//   short blocks (type 2):      slot s = 3 m + w holds elements (6m+w, 6m+3+w), i.e. lines 2m and
//                               2m+1 of window w (the [192][3] view of loop.c:1375-1376)
// Per-band quantities (xmin, distortion, scalefactors) are held "band per lane": band b lives in
// lane b&31 of register b>>5.  long: b = sfb (0..20); short: b = 3*sfb + window (0..35).
//
// Decisions are bit-exact restatements of the reference (same probe sequence, same tie rules); the
// only relaxation is the ORDER of FP64 additions inside per-band energy/noise sums (documented in
// changes a compare only if two doubles agree to ~1e-16).
#pragma once
#include "simt.h"

namespace mp3gpu {

using simt::PerThread;
using simt::WarpCtx;

// work statistics of the host emulation (tests / tools only): granule-channels, outer iterations, probes, quantiser rows,
// amplified bands, refreshed rows
#if !SIMT_DEV && defined(MP3GPU_RL_STATS)
extern long g_rl_stats[16];
#define RL_STAT(i, n) (g_rl_stats[i] += (n))
#else
#define RL_STAT(i, n) ((void)0)
#endif

// ---- constant tables (host-built with the reference's libm expressions; see tables.cpp) ---------
// Hot part: copied to shared memory by every CTA.
// glut[g][16 x + y]: code length + sign bits of the pair (x, y) in every candidate table of group g,
// three 10-bit fields, so ONE shared load per pair prices all sibling tables new_choose_table()
// compares (loop.c:1793-1900):
//   g0 {1}   g1 {2,3}   g2 {5,6}   g3 {7,8,9}   g4 {10,11,12}   g5 {13,15}
//   g6 {16..23, 24..31, #escapes}   g7 {15, 24, #escapes}   (x, y clamped to 15; linbits added per escape)
// A region priced with group g holds no value above gmax = 1, 2, 3, 5, 7, 14, 15, 15, so only the rows x <= gmax are stored
// (1120 instead of 2048 entries: the 3.6 KB buy the 28th warp of the CTA); group g starts at entry 16 * GLUT_OFF16(g).
#define GLUT_ENTRIES 1120
#define GLUT_OFF16(g) ((unsigned)((0x3626170f09050200ull >> (8 * (g))) & 0xff))
struct alignas(16) RateHot {
    unsigned int glut[GLUT_ENTRIES];
    unsigned int c1lut[16];         // count1 quad p: (hlen32 + signs) | (hlen33 + signs) << 16
    double pre1[4], pre2[4];        // pow(sqrt 2, n), pow(sqrt 2, 2n)  (loop.c:1205-1210)
    double ifqstep, ifqstep2;       // sqrt(2), sqrt(2)*sqrt(2)          (loop.c:1252,1293)
    double log2c, pad0;             // log(2.0)                           (loop.c:632)
    short sfb_l[24], sfb_s[16];     // Table B.8 for this sample rate
    unsigned short hlinmax[36];
    unsigned char hlinbits[36];
    unsigned char band_long[288];   // sfb of long pair s (21 = above last band)
    unsigned char band_short[288];  // 3*sfb + w of short slot s (>= 36 = above last band)
    unsigned char subdiv[290][2];   // region0/1_count for big_values (long blocks, loop.c:1596-1690)
    unsigned char pretab[24];
    unsigned char pad1[8];
    // band sums of long blocks (band_sums): the 21 scalefactor bands are cut into 32 runs of slots of near-equal length, one per
    // lane (bands wider than the run limit take 2..4 adjacent lanes); lane l sums slots [bs_start[l], bs_start[l] + bs_count[l])
    unsigned short bs_start[32];
    unsigned char bs_count[32];
    unsigned char bs_more[32];      // runs of the same band that follow this lane's run (the band's first lane adds them up)
    unsigned char bs_first[24];     // lane of the first run of band b
    unsigned char pad2[8];
};

// Cold part: stays in global memory (L1/L2 resident; touched rarely or with warp-uniform addresses)
struct alignas(16) RateTables {
    RateHot hot;
    double pow_nint_tab[2050];  // [p] = (p - 0.4054)^(4/3), p = 1..2048 (pow_nint.c:13-19); [0] unused
    double pow43[2048];         // p^(4/3) (loop.c:1017-1021)
    double step[512];           // 2^(q/4), q = -256..255 (loop.c:1020,1386); index q + 256
    double ostep[512];          // 1 / step
    float ostep34[512];         // (1 / step)^(3/4): scales the cached |xr|^(3/4) estimates (quantize_all)
    unsigned char hlen[1412];   // Table B.7 code lengths, flat (table builders / diagnostics)
    unsigned short hoff[34];
    unsigned char hxlen[34], hlinbits[34];
    unsigned short hlinmax[34];
    unsigned char pad[6];
};

struct GrInfoOut {  // 20 ints, same order as the reference dump (oracle GR_FIELDS)
    int part2_3_length, big_values, count1, global_gain, scalefac_compress;
    int window_switching_flag, block_type, mixed_block_flag;
    int table_select[3];
    int region0_count, region1_count, preflag, scalefac_scale, count1table_select;
    int part2_length, address1, address2, address3;
};

// per-stream state that survives between frames (reference statics: reservoir.c:36-37, loop.c:618-621)
struct LoopStreamState {
    int resv_size;
    int xrmax[4];      // [gr*2+ch]
    int en_tot[4];
    int addr[4][3];    // address1..3 of each (gr,ch): subdivide() leaves them stale when big_values == 0
    int pad[3];
};

// band-per-lane part of the state: calc_scfsi's `en` and `xm` statics, [gr*2+ch][sfb]
// (int-typed log2 energies, sic — loop.c:619-620)
struct LoopLaneState {
    int en[4][32];
    int xm[4][32];
};

struct FrameGeom {
    int n_ch, mean_bits, bits_per_frame;
    int mean_per_ch;   // mean_bits / n_ch           } derived once on the host (frame_geom_derive): a run-time integer division
    int resv_max;      // ResvFrameBegin's ResvMax   } is ~25 instructions, and this kernel pays for code size
    int xr_f32;        // the spectra the rate loop reads are float[576] per granule-channel (FP32 front-end variant), not double
};
inline void frame_geom_derive(FrameGeom *G)
{
    G->xr_f32 = 0;
    G->mean_per_ch = G->mean_bits / G->n_ch;
    G->resv_max = (G->bits_per_frame > 7680) ? 0 : ((7680 - G->bits_per_frame > 4088) ? 4088 : 7680 - G->bits_per_frame);   // reservoir.c:62-84
}

// ---- per-warp working set in shared memory ---------------------------------------------------------
// Slot order (k = 0..8, lane = 0..31, slot s = lane + 32 k in [0,288)):
//   long / start / stop blocks: slot s holds elements (2s, 2s+1)            -> Huffman pair s
//   short blocks (type 2):      slot s = 3 m + w holds elements (6m+w, 6m+3+w), i.e. lines 2m and
//                               2m+1 of window w (the [192][3] view of loop.c:1375-1376)
struct alignas(16) D2 { double x, y; };
struct alignas(4) U2 { unsigned short x, y; };
struct alignas(8) F2 { float x, y; };
struct alignas(16) RateWarpSmem {
    D2 xs[288];        // |xr| of the slot's two elements (amplified in place by the outer loop)
    double scr[288];   // per-slot energies / noise for the band sums; between refresh_pow34() and calc_noise it holds
                       // F2 ys[288] = FP32 estimates of |xr|^(3/4), the probe-invariant part of the quantiser
    U2 ix[288];        // quantised values of the slot
};

// ---- small helpers --------------------------------------------------------------------------------
SIMT_FN float pow075_estimate(float x)
{
#if SIMT_DEV
    // ex2(0.75 lg2 x) on the MUFU unit — what __powf() computes, without its denormal scaling (6 more instructions per
    // value): |xr| below 1.2e-38 gives the estimate 0, and such a line quantises to 0 at every step size.
    // relative error ~1.3e-6 typical, only ever used inside a verified margin
    float l, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
    l = __fmul_rn(l, 0.75f);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l));
    return r;
#else
    return powf(x, 0.75f);
#endif
}

// pow_nint(): largest p in [0,2047] with x >= tab[p] (tab strictly increasing, tab[0] := -inf) — the
// value the reference finds with gallop + binary search (pow_nint.h:16-50).  p is only a starting guess.
SIMT_NOINLINE int quant1_exact(const double *tab, double x, int p)
{
    if (p < 0) p = 0;
    if (p > 2047) p = 2047;
    while (p > 0 && x < tab[p]) p--;
    while (p < 2047 && x >= tab[p + 1]) p++;
    return p;
}

// libm calls that sit outside the hot loop are kept out of line: the kernel's working set of
// instructions must fit the instruction cache (see profiles/r01_a_baseline.md)
SIMT_NOINLINE double ref_log(double x) { return log(x); }
SIMT_NOINLINE double ref_exp(double x) { return exp(x); }
SIMT_NOINLINE double ref_div(double a, double b) { return a / b; }   // IEEE division: ~30 instructions per inlined copy

SIMT_FN int nint_ref(double in) { return (in < 0) ? (int)(in - 0.5) : (int)(in + 0.5); }

// The three small selectors below are nibble look-ups in 64-bit immediates instead of compare chains or table
// walks: the probe is the hot loop of the kernel and its code must stay resident in the 32 KB instruction cache.

// first table whose range covers max (loop.c:1813-1818 / 1921-1928): max in 1..14 -> 1,2,5,7,7,10,10,13..13
SIMT_FN int table_for_small_max(int max)
{
    return (int)((0xddddddddaa775210ull >> (4 * (max & 15))) & 15);
}

// first ESC table in [lo, hi) whose linmax covers m15 = max - 15 (loop.c:1873-1884, 1930-1941).
// linmax is 2^linbits - 1, so the choice only depends on the bit length n of m15:
//   tables 16..23 linbits 1,2,3,4,6,8,10,13   tables 24..31 linbits 4,5,6,7,8,9,11,13   (table 15: m15 == 0 only)
SIMT_FN int esc_table(const RateHot &H, int lo, int hi, int m15)
{
    (void)H; (void)hi;
    if (m15 <= 0 && lo == 15) return 15;
    const int n = m15 > 0 ? 32 - simt::clz((unsigned)m15) : 0;       // 0..13
    if (n > 13) return 0;
    return (lo >= 24) ? 24 + (int)((0x77665432100000ull >> (4 * n)) & 15)
                      : 16 + (int)((0x77766554432100ull >> (4 * n)) & 15);
}

// glut group of a region whose largest value is max (> 0): 1->0, 2->1, 3->2, 4..5->3, 6..7->4, 8..14->5, 15->7, >15->6
SIMT_FN int group_for_max(int max)
{
    return max > 15 ? 6 : (int)((0x7555555544332100ull >> (4 * max)) & 15);
}

struct BandRegs {   // band per lane: band b lives in lane b & 31 of register b >> 5
    PerThread<double> xmin[2], xfsf[2];
    PerThread<int> sf[2];
};

struct CountResult {
    int kz;     // rows (32 slots) >= kz of the ix[] working set hold zeros (carried from probe to probe, see quantize_all)
    int bits, big_values, count1, count1table_select;
    int region0_count, region1_count, address1, address2, address3;
    int table_select[3];
};

SIMT_FN int slot_e0(bool is_short, int s)
{
    if (!is_short) return 2 * s;
    int m = s / 3;
    return 6 * m + (s - 3 * m);
}

// quantize(): loop.c:1360-1428 with subblock_gain == 0 and mixed_block_flag == 0 (always, l3psy.c:739).
// x^(3/4) + 0.4054 is estimated in FP32 as t = |xr|^(3/4) * (1/step)^(3/4) + 0.4054 from the cached per-slot power.
// The estimate's relative error is < 2.2e-6 (two MUFU ops on |log2| <= 24, FP32 roundings); with the 5x-conservative
// margin m = 1e-5 t + 1e-6 the truncation is exact whenever floor(t - m) == floor(t + m) (and t is in range and a
// number), otherwise the FP64 table comparison on |xr| / step decides (pow_nint.h:16-50) — exact by construction.
// Also returns, per lane, the last slot holding a non-zero value and the last slot holding a value
// > 1 (calc_runlen's scan, loop.c:1498-1517) since the values are at hand.
// The loop is the hottest of the kernel (30 probes x 9 trips per granule-channel): one trip is ~35 instructions.
SIMT_FN void quantize_all(const WarpCtx &w, const RateTables &T, RateWarpSmem &M, int q, const PerThread<float> &rowmax, int &kz,
                          PerThread<int> &nzmax, PerThread<int> &bigmax)
{
    const int qi = (q > 255 ? 255 : q) + 256;
    const double ostep = T.ostep[qi];
    const float of = T.ostep34[qi];
    // rows whose largest estimate stays below 0.999 quantise to zero for certain (the estimate is good to 2.2e-6): only the
    // rows up to the last one that can hold a non-zero value are computed, the others are zero-filled if they are not yet
    PerThread<int> live;
    FOR_THREADS(w)
    live() = lane < 9 && !(simt::ffma(rowmax(), of, 0.4054f) < 0.999f);      // nan estimates count as live
    END_THREADS
    const unsigned lm = w.ballot(live);
    const int k_lim = 32 - simt::clz(lm);
    // the previous probe's count loops READ ix[] of other lanes' slots; their results went through warp reductions that every
    // lane waited for, so those reads are done — the barrier states the write-after-read order in the memory model's terms
    w.sync();
    RL_STAT(2, 1); RL_STAT(3, k_lim);
    FOR_THREADS(w)
    const F2 *ys = reinterpret_cast<const F2 *>(M.scr);
    unsigned *ixw = reinterpret_cast<unsigned *>(M.ix);
    int nz = -1, bg = -1;
#pragma unroll 1
    for (int s = lane; s < 32 * k_lim; s += 32) {
        const F2 y = ys[s];
        // clamped at 2040: an out-of-range (or nan) estimate sits exactly on an integer and therefore fails the test below
        const float ta = simt::fmin_(simt::ffma(y.x, of, 0.4054f), 2040.0f), tb = simt::fmin_(simt::ffma(y.y, of, 0.4054f), 2040.0f);
        // floor(t - m) and floor(t + m), m = 1e-5 t + 1e-6 (one fused multiply-add each), as "magic" floats 2^23 + floor(.)
        // (t - m >= 0.4, t + m < 2041): equal <=> accepted
        const float fa = simt::floor_magic(simt::ffma(ta, 1.0f - 1e-5f, -1e-6f)), fb = simt::floor_magic(simt::ffma(tb, 1.0f - 1e-5f, -1e-6f));
        const bool oka = (fa == simt::floor_magic(simt::ffma(ta, 1.0f + 1e-5f, 1e-6f)));
        const bool okb = (fb == simt::floor_magic(simt::ffma(tb, 1.0f + 1e-5f, 1e-6f)));
        unsigned pk = simt::pack_lo16(simt::fbits(fa), simt::fbits(fb));          // the values are the low mantissa bits: a | b << 16
        if (!(oka && okb)) {
            const D2 x = M.xs[s];
            unsigned a = pk & 0xffffu, b = pk >> 16;
            if (!oka) a = (unsigned)quant1_exact(T.pow_nint_tab, simt::dmul(x.x, ostep), (int)a);
            if (!okb) b = (unsigned)quant1_exact(T.pow_nint_tab, simt::dmul(x.y, ostep), (int)b);
            pk = a | (b << 16);
        }
        ixw[s] = pk;
        if (pk != 0) nz = s;
        if ((pk & 0xfffefffeu) != 0) bg = s;      // a > 1 || b > 1
    }
#pragma unroll 1
    for (int s = lane + 32 * k_lim; s < 32 * kz; s += 32) ixw[s] = 0u;
    nzmax() = nz; bigmax() = bg;
    END_THREADS
    kz = k_lim;
    w.sync();
}

// ys[s] = |xr|^(3/4) of the slot's two elements (FP32 estimate), into the scratch area: valid until calc_noise
// overwrites it.  Called once per outer-loop iteration (xs only changes between iterations).
SIMT_FN void refresh_pow34(const WarpCtx &w, RateWarpSmem &M, PerThread<float> &rowmax)
{
    F2 *ys = reinterpret_cast<F2 *>(M.scr);
    w.sync();
#pragma unroll 1
    for (int k = 0; k < 9; k++) {
        PerThread<float> m;
        FOR_THREADS(w)
        const int s = lane + 32 * k;
        const D2 x = M.xs[s];
        F2 y; y.x = pow075_estimate((float)x.x); y.y = pow075_estimate((float)x.y);
        ys[s] = y;
        m() = y.x > y.y ? y.x : y.y;
        END_THREADS
        const float r = w.reduce_max_nonneg(m);       // lane k keeps the largest estimate of row k (quantize_all)
        FOR_THREADS(w)
        if (lane == k) rowmax() = r;
        END_THREADS
    }
    w.sync();
}

SIMT_FN int count_regions(const WarpCtx &w, const RateHot &H, const RateWarpSmem &M, bool is_short, int bvr, int c1bits, CountResult &C);

// count_bits() = calc_runlen + count1_bitcount + subdivide + bigv_tab_select + bigv_bitcount
// (loop.c:2099-2113 and 590-594) on the quantised values in shared memory.  `C` carries address1..3
// across probes exactly like the reference's cod_info does (subdivide() leaves them untouched when
// big_values == 0, and bigv_tab_select()/bigv_bitcount() then still walk them, loop.c:1642-1647,1764-1772).
SIMT_FN int count_all(const WarpCtx &w, const RateHot &H, const RateWarpSmem &M, bool is_short, bool wsf,
                      const PerThread<int> &nzmax, const PerThread<int> &bigmax, CountResult &C)
{
    int c1bits = 0, bvr;
    C.table_select[0] = C.table_select[1] = C.table_select[2] = 0;
    if (is_short) {
        // loop.c:1492-1496, 1667-1674: fixed partition, lines < 12 (slots < 18) form region 0
        C.big_values = 288; C.count1 = 0; C.count1table_select = 1;
        C.region0_count = 8; C.region1_count = 36;
        C.address1 = 36; C.address2 = 576; C.address3 = 0;
        bvr = 576;
    } else {
        // ---- calc_runlen, loop.c:1498-1517 ----
        const int n = w.reduce_max(nzmax) + 1;
        const int B = w.reduce_max(bigmax);
        const int count1 = (n - 1 - B) >> 1;
        const int bv = n - 2 * count1;
        C.big_values = bv; C.count1 = count1;
        // ---- count1_bitcount, loop.c:1531-1590: quad t = slots (bv+2t, bv+2t+1) ----
        C.count1table_select = 1;
        if (count1 > 0) {
            PerThread<int> acc;
            FOR_THREADS(w)
            int a = 0;
#pragma unroll 1
            for (int t = lane; t < count1; t += 32) {
                const U2 u = M.ix[bv + 2 * t], v = M.ix[bv + 2 * t + 1];
                const int p = (u.x & 1) | ((u.y & 1) << 1) | ((v.x & 1) << 2) | ((v.y & 1) << 3);
                a += (int)H.c1lut[p];
            }
            acc() = a;
            END_THREADS
            const int both = w.reduce_add(acc);
            const int sum0 = both & 0xffff, sum1 = both >> 16;
            if (sum0 < sum1) { c1bits = sum0; C.count1table_select = 0; }
            else c1bits = sum1;
        }
        // ---- subdivide, loop.c:1638-1704 ----
        bvr = 2 * bv;
        if (bv == 0) {
            C.region0_count = 0; C.region1_count = 0;   // address1..3 stay stale
        } else if (!wsf) {
            C.region0_count = H.subdiv[bv][0];
            C.region1_count = H.subdiv[bv][1];
            C.address1 = H.sfb_l[C.region0_count + 1];
            C.address2 = H.sfb_l[C.region0_count + C.region1_count + 2];
            C.address3 = bvr;
        } else {
            C.region0_count = 7; C.region1_count = 13;
            C.address1 = H.sfb_l[8]; C.address2 = bvr; C.address3 = 0;
        }
    }
    return count_regions(w, H, M, is_short, bvr, c1bits, C);
}

// bigv_tab_select + bigv_bitcount (loop.c:1717-1775, 1793-1943, 1954-2016) on the regions given by C.address1 / C.address2
// and bvr = 2 * big_values; adds the bits to c1bits, sets C.table_select and C.bits.
// region of element e: R0 = [0,a1), R1 = [a1,a2) if a2 > a1, R2 = [a2,bvr) if bvr > a2.  (Region-2
// counting in the reference runs over [a2, a3): a3 == bvr for plain long blocks and 0 for
// start/stop/short blocks, where region 2 is never selected, so [a2,bvr) covers both.)
SIMT_FN int count_regions(const WarpCtx &w, const RateHot &H, const RateWarpSmem &M, bool is_short, int bvr, int c1bits, CountResult &C)
{
    const int a1 = C.address1, a2 = C.address2;
    const bool has0 = a1 > 0, has1 = a2 > a1, has2 = !is_short && bvr > a2;
    // Slots s hold elements 2 s, 2 s + 1 and a1, a2, bvr are even, so region r is the slot range [lo[r], hi[r]) (empty when
    // the region is absent; everything is capped at the 288 slots of the granule).  One loop per region, lanes striding
    // through it: a trip is a load and one packed max (or one table look-up and an add) instead of three-way selects.
    const int S1 = a1 >> 1, S2 = a2 >> 1, S3 = bvr >> 1;
    int lo[3], hi[3];
    lo[0] = 0;                  hi[0] = has0 ? (S1 < 288 ? S1 : 288) : 0;
    lo[1] = S1;                 hi[1] = has1 ? (S2 < 288 ? S2 : 288) : 0;
    lo[2] = S1 > S2 ? S1 : S2;  hi[2] = has2 ? (S3 < 288 ? S3 : 288) : 0;
    const unsigned *ixw = reinterpret_cast<const unsigned *>(M.ix);      // x | y << 16
    int grp[3] = {-1, -1, -1}, rmax[3] = {0, 0, 0};
    bool any = false;
    // the three per-lane maxima first, then the three warp reductions back to back: they are independent, so their latencies
    // overlap (the kernel is bound by the serial chain of a probe, not by issue slots: -4.7 % time for the same instructions)
    PerThread<int> mx[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        FOR_THREADS(w)
        unsigned m = 0;
#pragma unroll 1
        for (int sl = lo[r] + lane; sl < hi[r]; sl += 32) m = simt::vmaxu2(m, ixw[sl]);   // per-halfword maximum
        mx[r]() = (int)simt::umax(m & 0xffffu, m >> 16);
        END_THREADS
    }
#pragma unroll
    for (int r = 0; r < 3; r++) {
        rmax[r] = w.reduce_max(mx[r]);
        if (rmax[r] > 0) { grp[r] = group_for_max(rmax[r]); any = true; }
    }
    int bits = c1bits;
    if (any) {
        const unsigned *gl = &H.glut[0];
#pragma unroll
        for (int r = 0; r < 3; r++) {
            if (grp[r] < 0) continue;
            const int g = grp[r], max = rmax[r];
            PerThread<int> lo_, hi_;
            FOR_THREADS(w)
            const unsigned *gt = gl + 16 * GLUT_OFF16(g);
            unsigned acc = 0;
#pragma unroll 1
            for (int sl = lo[r] + lane; sl < hi[r]; sl += 32) {
                const unsigned c = simt::vminu2(ixw[sl], 0x000f000fu);              // min(x, 15) | min(y, 15) << 16
                acc += gt[((c << 4) | (c >> 16)) & 0xffu];                            // 16 x + y
            }
            // per lane <= 9 pairs x <= 63 per field: no carry between the 10-bit fields
            lo_() = (int)((acc & 1023u) | (((acc >> 10) & 1023u) << 16)); hi_() = (int)(acc >> 20);
            END_THREADS
            // both reductions back to back, needed or not (the third field is read for groups 3, 4, 6, 7 only): a reduction
            // issued behind a branch on the first one's result would wait for it
            const int both = w.reduce_add(lo_), third = w.reduce_add(hi_);
            int s0 = both & 0xffff, s1 = both >> 16, choice;
            if (is_short) {
                // choose_table (by max alone), loop.c:1908-1943: first table covering max; 15 -> table 15
                if (g == 7) { choice = 15; }                               // glut g7 field 0 = table 15
                else if (g == 6) { choice = esc_table(H, 16, 24, max - 15); s0 += (int)H.hlinbits[choice] * third; }
                else choice = table_for_small_max(max);
            } else if (g >= 6) {
                // ESC pair, strict '<' (loop.c:1870-1898)
                const int nesc = third;
                const int c0 = esc_table(H, 15, 24, max - 15), c1 = esc_table(H, 24, 32, max - 15);
                s0 += (int)H.hlinbits[c0] * nesc;
                s1 += (int)H.hlinbits[c1] * nesc;
                choice = c0;
                if (s1 < s0) { choice = c1; s0 = s1; }
            } else {
                // sibling tables, '<=' chains (loop.c:1826-1868)
                choice = table_for_small_max(max);
                if (g >= 1) {
                    const int c1 = (g == 1) ? 3 : (g == 2) ? 6 : (g == 3) ? 8 : (g == 4) ? 11 : 15;
                    if (s1 <= s0) { choice = c1; s0 = s1; }
                    if (g == 3 || g == 4) {
                        const int s2 = third;
                        if (s2 <= s0) { choice = (g == 3) ? 9 : 12; s0 = s2; }
                    }
                }
            }
            C.table_select[r] = choice;
            bits += s0;
        }
    }
    C.bits = bits;
    return bits;
}

// one quantize + count_bits probe at step q
SIMT_FN int probe(const WarpCtx &w, const RateHot &H, const RateTables &T, RateWarpSmem &M, bool is_short, bool wsf, int q,
                  const PerThread<float> &rowmax, CountResult &C)
{
    PerThread<int> nzmax, bigmax;
    quantize_all(w, T, M, q, rowmax, C.kz, nzmax, bigmax);
    return count_all(w, H, M, is_short, wsf, nzmax, bigmax, C);
}

// sum over the slots of each band of the per-slot values in scr[288]; result band-per-lane.
// long: band b = pairs [sfb_l[b]/2, sfb_l[b+1]/2), b < 21; short: b = 3 sfb + w -> slots 3m+w, m in
// [sfb_s[sfb]/2, sfb_s[sfb+1]/2), sfb < 12.
SIMT_FN double band_sum_one(const RateHot &T, const double *scr, bool is_short, int b)
{
    // slots of band b: start, stride, count (two interleaved accumulators, even/odd slot of the band)
    int st, stride, n;
    if (!is_short) {
        if (b >= 21) return 0.0;
        st = T.sfb_l[b] >> 1; stride = 1; n = (T.sfb_l[b + 1] >> 1) - st;
    } else {
        if (b >= 36) return 0.0;
        const int sfb = b / 3, wi = b - 3 * sfb;
        const int lo = T.sfb_s[sfb] >> 1;
        st = 3 * lo + wi; stride = 3; n = (T.sfb_s[sfb + 1] >> 1) - lo;
    }
    double a0 = 0.0, a1 = 0.0;
    const double *p = scr + st;
    int i = 0;
#pragma unroll 1
    for (; i + 1 < n; i += 2) { a0 = simt::dadd(a0, p[0]); a1 = simt::dadd(a1, p[stride]); p += 2 * stride; }
    if (i < n) a0 = simt::dadd(a0, p[0]);
    return simt::dadd(a0, a1);
}

SIMT_FN void band_sums(const WarpCtx &w, const RateHot &T, const double *scr, bool is_short, PerThread<double> out[2])
{
    if (is_short) {
        FOR_THREADS(w)
        out[0]() = band_sum_one(T, scr, true, lane);
        out[1]() = band_sum_one(T, scr, true, lane + 32);
        END_THREADS
        return;
    }
    // long blocks: 21 bands of 2..38 slots.  One band per lane would leave a third of the lanes idle and the warp waiting for
    // the widest band; the bands are cut into 32 runs of at most ~11 slots instead (RateHot::bs_*), the partial sums of a band
    // meet in its first lane (in run order) and are handed to lane b.
    PerThread<double> part, d1, d2, d3, tot;
    PerThread<int> src;
    FOR_THREADS(w)
    {
        const double *p = scr + T.bs_start[lane];
        const int n = T.bs_count[lane];
        double a0 = 0.0, a1 = 0.0;
        int i = 0;
#pragma unroll 1
        for (; i + 1 < n; i += 2) { a0 = simt::dadd(a0, p[0]); a1 = simt::dadd(a1, p[1]); p += 2; }
        if (i < n) a0 = simt::dadd(a0, p[0]);
        part() = simt::dadd(a0, a1);
    }
    END_THREADS
    w.shfl_down(d1, part, 1); w.shfl_down(d2, part, 2); w.shfl_down(d3, part, 3);
    FOR_THREADS(w)
    {
        const int more = T.bs_more[lane];
        double v = part();
        if (more >= 1) v = simt::dadd(v, d1());
        if (more >= 2) v = simt::dadd(v, d2());
        if (more >= 3) v = simt::dadd(v, d3());
        tot() = v;
        src() = lane < 21 ? T.bs_first[lane] : 0;
    }
    END_THREADS
    w.shfl_idx(out[0], tot, src);
    FOR_THREADS(w)
    if (lane >= 21) out[0]() = 0.0;
    out[1]() = 0.0;
    END_THREADS
}

SIMT_FN double band_width(const RateHot &T, bool is_short, int b)
{
    if (!is_short) return (double)(T.sfb_l[b + 1] - T.sfb_l[b]);
    int sfb = b / 3;
    return (double)(T.sfb_s[sfb + 1] - T.sfb_s[sfb]);
}

SIMT_FN int part2_length_of(bool is_short, int gr, int compress, const int scfsi[4])
{
    // loop.c:731-784 (MPEG-1)
    const int s1t = (0x4433322211130000ull >> (4 * compress)) & 15;  // slen1_tab
    const int s2t = (0x3232132132103210ull >> (4 * compress)) & 15;  // slen2_tab
    if (is_short) return 18 * s1t + 18 * s2t;
    int bits = 0;
    if (gr == 0 || scfsi[0] == 0) bits += 6 * s1t;
    if (gr == 0 || scfsi[1] == 0) bits += 5 * s1t;
    if (gr == 0 || scfsi[2] == 0) bits += 5 * s2t;
    if (gr == 0 || scfsi[3] == 0) bits += 5 * s2t;
    return bits;
}

// ResvMaxBits, reservoir.c:101-134: the bits a granule-channel may use, from the reservoir level and its perceptual entropy
SIMT_FN int resv_max_bits(const FrameGeom &G, int resv_size, double pe)
{
    const int mean = G.mean_per_ch;
    const int resv_max = G.resv_max;
    int max_bits = mean > 4095 ? 4095 : mean;
    if (resv_max != 0) {
        int more_bits = (int)(pe * 3.1 - mean), add_bits = 0;
        if (more_bits > 100) {
            int frac = (resv_size * 6) / 10;
            add_bits = frac < more_bits ? frac : more_bits;
        }
        int over_bits = resv_size - ((resv_max * 8) / 10) - add_bits;
        if (over_bits > 0) add_bits += over_bits;
        max_bits += add_bits;
        if (max_bits > 4095) max_bits = 4095;
    }
    return max_bits;
}

struct Gr0Carry {             // what granule 1 may need from granule 0 of the same channel
    PerThread<int> sf0;       // long scalefactors of gr 0 (band per lane)
    int preflag, scalefac_scale;
};

// Encode one granule-channel.  xr: 576 doubles (mdct_sub output).  ratio: 21 (long) or 36 ([sfb][win])
// doubles.  Writes ix (signed, sign of xr applied as l3bitstream.c:115-125 does), gi, scalefac bytes.
// Returns part2_3_length (before ResvFrameEnd stuffing).
SIMT_FN int encode_gc(const WarpCtx &w, const RateHot &H, const RateTables &T, RateWarpSmem &M, const FrameGeom &G, LoopStreamState &S,
                      PerThread<int> st_en[4], PerThread<int> st_xm[4],
                      int gr, int ch, const double *xr, const double *ratio_l, const double *ratio_s, double pe,
                      int block_type, int scfsi[4], Gr0Carry &g0, short *ix_out, GrInfoOut *gi_out, unsigned char *sf_out,
                      int *max_bits_out)
{
    RL_STAT(0, 1);
    const bool is_short = (block_type == 2);
    const bool wsf = (block_type != 0);
    const int nb_l = is_short ? 0 : 21;   // sfb_lmax (gr_deco, loop.c:2063-2081)
    double *scr = M.scr;
    BandRegs Bd;
    PerThread<int> sign;      // bit 2k / 2k+1: sign of the slot's elements
    PerThread<double> t0, t1;

    // ---- load xr into registers -----------------------------------------------------------------
    FOR_THREADS(w)
    int sg = 0;
    double mx = 0.0, e2 = 0.0, lg = 0.0, pr = 1.0;
#pragma unroll 3
    for (int k = 0; k < 9; k++) {
        int s = lane + 32 * k;
        int e0 = slot_e0(is_short, s);
        int e1 = is_short ? e0 + 3 : e0 + 1;
        double a, b;
        if (G.xr_f32) { a = (double)reinterpret_cast<const float *>(xr)[e0]; b = (double)reinterpret_cast<const float *>(xr)[e1]; }
        else { a = xr[e0]; b = xr[e1]; }
        if (a < 0) sg |= 1 << (2 * k);
        if (b < 0) sg |= 2 << (2 * k);
        a = fabs(a); b = fabs(b);
        D2 x2; x2.x = a; x2.y = b;
        M.xs[s] = x2;
        U2 z; z.x = 0; z.y = 0;
        M.ix[s] = z;
        mx = fmax(mx, fmax(a, b));
        double a2 = simt::dmul(a, a), b2 = simt::dmul(b, b);
        scr[s] = simt::dadd(a2, b2);
        e2 = simt::dadd(e2, simt::dadd(a2, b2));
        // quantanf_init's sum of log(xr^2) over the non-zero lines (loop.c:380-386): one log per six lines on the product of
        // their squares; like the lane-wise order of the sum this moves sum1 by a few ulp, which nint(8 ln(sfm)) only sees on
        // an exact rounding boundary.  Six squares stay normal numbers for |xr| in 1e-25 .. 1e+25; rounding residue of a
        // cancelling MDCT can be smaller, so the product is flushed early whenever it leaves 1e-200 .. 1e+200 (a square
        // itself cannot underflow before |xr| < 1e-154, where the reference's own xr * xr is already denormal).
        pr = simt::dmul(pr, a != 0 ? a2 : 1.0);
        pr = simt::dmul(pr, b != 0 ? b2 : 1.0);
        if (k % 3 == 2 || pr < 1e-200 || pr > 1e200) { lg = simt::dadd(lg, ref_log(pr)); pr = 1.0; }
    }
    sign() = sg; t0() = mx; t1() = e2;
    Bd.xfsf[0]() = lg;  // borrowed as scratch for the log sum
    END_THREADS
    w.sync();
    const double xrmax = w.reduce_max(t0);
    const double sum2 = w.reduce_add(t1);
    const double sum1 = w.reduce_add(Bd.xfsf[0]);

    // ---- calc_xmin, loop.c:1085-1119 ------------------------------------------------------------
    PerThread<double> en[2];
    band_sums(w, H, scr, is_short, en);
    FOR_THREADS(w)
#pragma unroll
    for (int h = 0; h < 2; h++) {
        int b = lane + 32 * h;
        double v = 0.0;
        if (!is_short) { if (b < 21) v = ref_div(simt::dmul(ratio_l[b], en[h]()), band_width(H, false, b)); }
        else if (b < 36) v = ref_div(simt::dmul(ratio_s[b], en[h]()), band_width(H, true, b));
        Bd.xmin[h]() = v;
        Bd.xfsf[h]() = 0.0;
        Bd.sf[h]() = 0;
    }
    END_THREADS

    // ---- calc_scfsi, loop.c:615-722 (int-typed statics and swapped [ch][gr] indices kept) ----------
    {
        const int cur = gr * 2 + ch;
        S.xrmax[cur] = (int)xrmax;
        S.en_tot[cur] = (sum2 == 0.0) ? 0 : (int)ref_div(ref_log(sum2), H.log2c);
        if (!is_short) {
            FOR_THREADS(w)
            if (lane < 21) {
                double e = en[0](), xm = Bd.xmin[0]();
                int ev = (e == 0.0) ? 0 : (int)ref_div(ref_log(e), H.log2c);
                int xv = (xm == 0.0) ? 0 : (int)ref_div(ref_log(xm), H.log2c);
#pragma unroll
                for (int i = 0; i < 4; i++) if (i == cur) { st_en[i]() = ev; st_xm[i]() = xv; }
            }
            END_THREADS
        }
        if (gr == 1) {
            int condition = 0;
            for (int gr2 = 0; gr2 < 2; gr2++) {
                if (S.xrmax[ch * 2 + gr2] != 0) condition++;
                if (!is_short) condition++;
            }
            condition++;  // loop.c:683 compares a pointer difference (== 2) against 10: always true
            PerThread<int> d_en, d_xm;
            FOR_THREADS(w)
            int a = 0, b = 0, c = 0, d = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (i == ch * 2) { a = st_en[i](); c = st_xm[i](); }
                if (i == ch * 2 + 1) { b = st_en[i](); d = st_xm[i](); }
            }
            d_en() = (lane < 21) ? (a > b ? a - b : b - a) : 0;
            d_xm() = (lane < 21) ? (c > d ? c - d : d - c) : 0;
            END_THREADS
            if (w.reduce_add(d_en) < 100) condition++;
            if (condition == 6) {
#pragma unroll 1
                for (int band = 0; band < 4; band++) {
                    const int lo = (band == 0) ? 0 : (band == 1) ? 6 : (band == 2) ? 11 : 16;
                    const int hi = (band == 0) ? 6 : (band == 1) ? 11 : (band == 2) ? 16 : 21;
                    PerThread<int> p0, p1;
                    FOR_THREADS(w)
                    bool in = lane >= lo && lane < hi;
                    p0() = in ? d_en() : 0; p1() = in ? d_xm() : 0;
                    END_THREADS
                    scfsi[band] = (w.reduce_add(p0) < 10 && w.reduce_add(p1) < 10) ? 1 : 0;
                }
            } else scfsi[0] = scfsi[1] = scfsi[2] = scfsi[3] = 0;
        }
    }

    // ---- ResvMaxBits, reservoir.c:101-134 -------------------------------------------------------
    const int max_bits = resv_max_bits(G, S.resv_size, pe);
    if (max_bits_out) *max_bits_out = max_bits;

    // ---- iteration variables, loop.c:319-346 ------------------------------------------------------
    CountResult C;
    C.kz = 0;   // ix[] was cleared above
    PerThread<float> rowmax;
    C.bits = 0; C.big_values = 0; C.count1 = 0; C.count1table_select = 0;
    C.region0_count = C.region1_count = 0;
    C.address1 = S.addr[gr * 2 + ch][0]; C.address2 = S.addr[gr * 2 + ch][1]; C.address3 = S.addr[gr * 2 + ch][2];
    C.table_select[0] = C.table_select[1] = C.table_select[2] = 0;
    int preflag = 0, compress = 0, part2 = 0, part23 = 0, q = 0, bits = 0;

    if (xrmax != 0.0) {
        // quantanf_init, loop.c:369-402
        {
            int tp = 0;
            if (sum2 != 0.0) {
                double sfm = ref_div(ref_exp(ref_div(sum1, 576.0)), ref_div(sum2, 576.0));
                tp = nint_ref(8.0 * ref_log(sfm));
                if (tp < -100) tp = -100;
            }
            q = tp - 70;
        }
        // outer_loop, loop.c:415-558
        int save_preflag = 0, save_compress = 0, iteration = 0, status, over;
        PerThread<int> save_sf[2];
        do {
            iteration++;
            RL_STAT(1, 1);
            refresh_pow34(w, M, rowmax);
            part2 = part2_length_of(is_short, gr, compress, scfsi);
            const int huff_bits = max_bits - part2;
            // bin_search_StepSize(max_bits, ...) on the first iteration (loop.c:2119-2140), then inner_loop
            // (loop.c:569-606), as ONE probe site (code size: the probe is the hot loop of the kernel).
            // inner_loop starts by re-quantising at the step the binary search stopped at; that probe is
            // a pure function of (xr, q, stale addresses) and therefore equal to the one just made, so
            // its result is reused instead of recomputed.
            {
                bool searching = (iteration == 1);
                int top = q, bot = 200, next = q, last = q;
                for (;;) {
                    if (searching) { last = next; next = (top + bot) / 2; q = next; }  // aint((top+bot)/2.0)
                    bits = probe(w, H, T, M, is_short, wsf, q, rowmax, C);
                    if (searching) {
                        if (bits > max_bits) top = next; else bot = next;
                        if (bits != max_bits && (last - next > 1 || next - last > 1)) continue;
                        searching = false;
                    }
                    if (!(bits > huff_bits && q < 1024)) break;  // q guard: the reference assert()s huff_bits >= 0 (loop.c:579)
                    q += 1;
                }
            }

            // calc_noise, loop.c:1007-1069
            {
                const double step = T.step[(q > 255 ? 255 : q) + 256];
                FOR_THREADS(w)
#pragma unroll 3
                for (int k = 0; k < 9; k++) {
                    const D2 x = M.xs[lane + 32 * k];
                    const U2 v = M.ix[lane + 32 * k];
                    double da = simt::dsub(x.x, simt::dmul(T.pow43[v.x], step));
                    double db = simt::dsub(x.y, simt::dmul(T.pow43[v.y], step));
                    scr[lane + 32 * k] = simt::dadd(simt::dmul(da, da), simt::dmul(db, db));
                }
                END_THREADS
                w.sync();
                PerThread<double> ns[2];
                band_sums(w, H, scr, is_short, ns);
                FOR_THREADS(w)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    int b = lane + 32 * h;
                    bool valid = is_short ? (b < 36) : (b < 21);
                    Bd.xfsf[h]() = valid ? ref_div(ns[h](), band_width(H, is_short, b)) : 0.0;
                    save_sf[h]() = Bd.sf[h]();
                }
                END_THREADS
                w.sync();
            }
            save_preflag = preflag;
            save_compress = compress;

            // over-threshold mask (xfsf > xmin), band per lane -> 64-bit uniform mask
            PerThread<int> ov0, ov1;
            FOR_THREADS(w)
            ov0() = Bd.xfsf[0]() > Bd.xmin[0]();
            ov1() = Bd.xfsf[1]() > Bd.xmin[1]();
            END_THREADS
            unsigned m0 = w.ballot(ov0), m1 = w.ballot(ov1);
            if (!is_short) { m0 &= (1u << 21) - 1; m1 = 0; } else m1 &= 15u;

            // preemphasis, loop.c:1161-1216
            bool scfsi_any = (gr == 1) && (scfsi[0] | scfsi[1] | scfsi[2] | scfsi[3]);
            if (scfsi_any) preflag = g0.preflag;
            else if (block_type != 2 && preflag == 0 && ((m0 >> 17) & 15u) == 15u) {
                preflag = 1;
                FOR_THREADS(w)
                if (lane < nb_l) Bd.xmin[0]() = simt::dmul(Bd.xmin[0](), H.pre2[H.pretab[lane]]);
#pragma unroll 1
                for (int k = 0; k < 9; k++) {
                    const int s = lane + 32 * k;
                    const int b = H.band_long[s];        // preemphasis never runs for short blocks
                    if (b < nb_l) {
                        const double f = H.pre1[H.pretab[b]];
                        D2 x = M.xs[s];
                        x.x = simt::dmul(x.x, f); x.y = simt::dmul(x.y, f);
                        M.xs[s] = x;
                    }
                }
                END_THREADS
                // NB the reference re-tests xfsf > xmin in amp_scalefac_bands AFTER xmin was scaled
                FOR_THREADS(w)
                ov0() = Bd.xfsf[0]() > Bd.xmin[0]();
                END_THREADS
                m0 = w.ballot(ov0) & ((1u << 21) - 1);
            }

            // amp_scalefac_bands, loop.c:1225-1349
            {
                unsigned skip = 0;  // long bands frozen by scfsi
                bool copySF = false;
                if (scfsi_any) {
                    if (iteration == 1) copySF = true;
                    for (int band = 0; band < 4; band++)
                        if (scfsi[band]) {
                            const int lo = (band == 0) ? 0 : (band == 1) ? 6 : (band == 2) ? 11 : 16;
                            const int hi = (band == 0) ? 6 : (band == 1) ? 11 : (band == 2) ? 16 : 21;
                            skip |= ((1u << hi) - 1) & ~((1u << lo) - 1);
                        }
                    if (is_short) skip = 0;  // loop over sfb < sfb_lmax == 0 never runs
                }
                const unsigned amp0 = m0 & ~skip, amp1 = m1;
                over = simt::popc(amp0) + simt::popc(amp1);
                RL_STAT(4, over);
                const unsigned long long amp = (unsigned long long)amp0 | ((unsigned long long)amp1 << 32);
                FOR_THREADS(w)
                if (copySF && !is_short && ((skip >> lane) & 1)) Bd.sf[0]() = g0.sf0();
                if ((amp0 >> lane) & 1) { Bd.xmin[0]() = simt::dmul(Bd.xmin[0](), H.ifqstep2); Bd.sf[0]()++; }
                if ((amp1 >> lane) & 1) { Bd.xmin[1]() = simt::dmul(Bd.xmin[1](), H.ifqstep2); Bd.sf[1]()++; }
                if (amp != 0) {
#pragma unroll 1
                    for (int k = 0; k < 9; k++) {
                        const int s = lane + 32 * k;
                        const int b = is_short ? H.band_short[s] : H.band_long[s];
                        if (b < 36 && ((amp >> b) & 1)) {
                            D2 x = M.xs[s];
                            x.x = simt::dmul(x.x, H.ifqstep); x.y = simt::dmul(x.y, H.ifqstep);
                            M.xs[s] = x;
                        }
                    }
                }
                END_THREADS
            }
            // loop_break, loop.c:1131-1150; scale_bitcount, loop.c:792-857
            {
                PerThread<int> z0, z1, s1m, s2m;
                FOR_THREADS(w)
                bool v0 = is_short ? true : (lane < 21);
                bool v1 = is_short ? (lane < 4) : false;
                z0() = v0 && Bd.sf[0]() == 0;
                z1() = v1 && Bd.sf[1]() == 0;
                int a = 0, b = 0;
                if (!is_short) {
                    if (lane < 11) a = Bd.sf[0](); else if (lane < 21) b = Bd.sf[0]();
                } else {  // bands 3*sfb+w: sfb < 6 <=> b < 18
                    if (lane < 18) a = Bd.sf[0](); else b = Bd.sf[0]();
                    if (lane < 4) b = b > Bd.sf[1]() ? b : Bd.sf[1]();
                }
                s1m() = a; s2m() = b;
                END_THREADS
                status = (w.ballot(z0) | w.ballot(z1)) ? 0 : 1;
                if (status == 0) {
                    const int ms1 = w.reduce_max(s1m), ms2 = w.reduce_max(s2m);
                    // first scalefac_compress k whose (slen1, slen2) hold n1 = bitlength(ms1), n2 = bitlength(ms2) bits: the
                    // reference's walk over the 16 candidates as a look-up on (n1, n2), nibble 4 n1 + n2
                    const int n1 = 32 - simt::clz((unsigned)ms1), n2 = 32 - simt::clz((unsigned)ms2);
                    int ep = 2;
                    if (n1 <= 4 && n2 <= 3) {
                        ep = 0;
                        compress = n1 < 4 ? (int)((0xdcb4a98476543210ull >> (4 * (4 * n1 + n2))) & 15) : (int)((0xfeeeu >> (4 * n2)) & 15);
                    }
                    status = ep;
                }
            }
        } while (status == 0 && over > 0);
        preflag = save_preflag;
        compress = save_compress;
        FOR_THREADS(w)
        Bd.sf[0]() = save_sf[0]();
        Bd.sf[1]() = save_sf[1]();
        END_THREADS
        part2 = part2_length_of(is_short, gr, compress, scfsi);
        part23 = part2 + bits;
    }

    // ---- ResvAdjust + global_gain, loop.c:355-358 -------------------------------------------------
    S.resv_size += G.mean_per_ch - part23;
    FOR_THREADS(w)
    if (lane == 0) {   // part2_3_length is rewritten after ResvFrameEnd (stuffing bits)
        GrInfoOut gi;
        gi.part2_3_length = part23;
        gi.big_values = C.big_values; gi.count1 = C.count1;
        gi.global_gain = nint_ref((double)q + 210.0);
        gi.scalefac_compress = compress;
        gi.window_switching_flag = wsf; gi.block_type = block_type; gi.mixed_block_flag = 0;
        gi.table_select[0] = C.table_select[0]; gi.table_select[1] = C.table_select[1]; gi.table_select[2] = C.table_select[2];
        gi.region0_count = C.region0_count; gi.region1_count = C.region1_count;
        gi.preflag = preflag; gi.scalefac_scale = 0; gi.count1table_select = C.count1table_select;
        gi.part2_length = part2; gi.address1 = C.address1; gi.address2 = C.address2; gi.address3 = C.address3;
        *gi_out = gi;
    }
    END_THREADS
    S.addr[gr * 2 + ch][0] = C.address1; S.addr[gr * 2 + ch][1] = C.address2; S.addr[gr * 2 + ch][2] = C.address3;

    // ---- outputs ------------------------------------------------------------------------------------
    FOR_THREADS(w)
#pragma unroll 1
    for (int k = 0; k < 9; k++) {
        int s = lane + 32 * k;
        int e0 = slot_e0(is_short, s);
        int e1 = is_short ? e0 + 3 : e0 + 1;
        const U2 v = M.ix[s];
        int a = v.x, b = v.y;
        if ((sign() >> (2 * k)) & 1) a = -a;
        if ((sign() >> (2 * k + 1)) & 1) b = -b;
        ix_out[e0] = (short)a;
        ix_out[e1] = (short)b;
    }
    if (!is_short) { if (lane < 22) sf_out[lane] = (unsigned char)((lane < 21) ? Bd.sf[0]() : 0); }
    else { sf_out[lane] = (unsigned char)Bd.sf[0](); if (lane < 4) sf_out[32 + lane] = (unsigned char)Bd.sf[1](); }
    END_THREADS
    if (gr == 0) {
        FOR_THREADS(w)
        g0.sf0() = is_short ? 0 : Bd.sf[0]();
        END_THREADS
        g0.preflag = preflag; g0.scalefac_scale = 0;
    }
    w.sync();
    return part23;
}

// ResvFrameEnd, reservoir.c:155-226: trims the reservoir and pushes stuffing bits into part2_3_length.
SIMT_FN void resv_frame_end(const FrameGeom &G, LoopStreamState &S, int p23[4], int *resv_drain)
{
    const int resv_max = G.resv_max;
    if (G.n_ch == 2 && (G.mean_bits & 1)) S.resv_size += 1;
    int over_bits = S.resv_size - resv_max;
    if (over_bits < 0) over_bits = 0;
    S.resv_size -= over_bits;
    int stuffing = over_bits;
    if ((over_bits = S.resv_size % 8)) { stuffing += over_bits; S.resv_size -= over_bits; }
    *resv_drain = 0;
    if (stuffing) {
        if (p23[0] + stuffing < 4095) p23[0] += stuffing;
        else {
            for (int gr = 0; gr < 2; gr++)
                for (int ch = 0; ch < G.n_ch; ch++) {
                    if (stuffing == 0) break;
                    int extra = 4095 - p23[gr * 2 + ch];
                    int take = extra < stuffing ? extra : stuffing;
                    p23[gr * 2 + ch] += take;
                    stuffing -= take;
                }
            *resv_drain = stuffing;
        }
    }
}

}  // namespace mp3gpu

// ---------------------------------------------------------------------------------------------------
// stream driver: one warp, one stream, n_frames frames in order
// ---------------------------------------------------------------------------------------------------
namespace mp3gpu {

// what the psychoacoustic scan hands to the rate loop for one granule-channel
// (L3psycho_anal outputs: ratio_d[21], ratio_ds[12][3], *pe, cod_info->block_type; l3psy.h:32-34)
struct PsyOut {
    double pe;
    double ratio_l[21];
    double ratio_s[36];  // [sfb][window]
    int block_type;
    int pad;
};

struct FrameOut {
    int resv_drain;          // l3_side->resvDrain
    int main_data_begin;     // back pointer of THIS frame in bytes = reservoir size before the frame / 8
    unsigned char scfsi[2][4];
};

// gc index inside a stream chunk: g = (frame*2 + gr)*n_ch + ch
SIMT_FN void rate_loop_stream(const WarpCtx &w, const RateHot &H, const RateTables &T, RateWarpSmem &M, const FrameGeom &G, LoopStreamState &S,
                              PerThread<int> st_en[4], PerThread<int> st_xm[4], int n_frames,
                              const double *xr, const PsyOut *psy, short *ix, GrInfoOut *gi, unsigned char *sf,
                              FrameOut *fo, int *max_bits_dbg, int *p23_pre = nullptr)
{
    for (int f = 0; f < n_frames; f++) {
        int scfsi[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
        int p23[4] = {0, 0, 0, 0};
        Gr0Carry g0[2];
        g0[0].preflag = g0[1].preflag = 0;
        g0[0].scalefac_scale = g0[1].scalefac_scale = 0;
        const int mdb = S.resv_size / 8;
        for (int gr = 0; gr < 2; gr++)
            for (int ch = 0; ch < G.n_ch; ch++) {
                const int g = (f * 2 + gr) * G.n_ch + ch;
                const PsyOut &po = psy[g];
                const double *xg = G.xr_f32 ? reinterpret_cast<const double *>(reinterpret_cast<const float *>(xr) + (size_t)g * 576)
                                            : xr + (size_t)g * 576;
                p23[gr * 2 + ch] = encode_gc(w, H, T, M, G, S, st_en, st_xm, gr, ch, xg, po.ratio_l, po.ratio_s,
                                             po.pe, po.block_type, scfsi[ch], g0[ch], ix + (size_t)g * 576, gi + g,
                                             sf + (size_t)g * 40, max_bits_dbg ? max_bits_dbg + g : nullptr);
            }
        if (p23_pre) {                 // part2_3_length of the frame's granule-channels before the stuffing bits of ResvFrameEnd
            FOR_THREADS(w)
            if (lane == 0) { p23_pre[4 * f] = p23[0]; p23_pre[4 * f + 1] = p23[1]; p23_pre[4 * f + 2] = p23[2]; p23_pre[4 * f + 3] = p23[3]; }
            END_THREADS
        }
        int drain = 0;
        resv_frame_end(G, S, p23, &drain);
        FOR_THREADS(w)
        if (lane == 0) {
            for (int gr = 0; gr < 2; gr++)
                for (int ch = 0; ch < G.n_ch; ch++) {
                    const int g = (f * 2 + gr) * G.n_ch + ch;
                    gi[g].part2_3_length = p23[gr * 2 + ch];
                }
            fo[f].resv_drain = drain;
            fo[f].main_data_begin = mdb;
            for (int ch = 0; ch < 2; ch++)
                for (int b = 0; b < 4; b++) fo[f].scfsi[ch][b] = (unsigned char)scfsi[ch][b];
        }
        END_THREADS
    }
}

}  // namespace mp3gpu
