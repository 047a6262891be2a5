// tables.cpp — host-side builders for the constant tables (see tables.h).
// Every expression that the reference evaluates with libm at start-up is evaluated here with the
// same libm, in the same association order, so that the uploaded tables are bit-identical to the
// reference's (citations: /root/reference/src/<file>:<line>).
#include "tables.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <vector>

#include "iso_tables.h"

namespace mp3gpu {

static const double kRefPi = 3.14159265358979;   // common.h:199 — the reference's PI, not M_PI
static const double kLnToLog10 = 0.2302585093;   // common.h:204
static const double kTwoPi = 6.28318530717958647692;  // subs.c:25

int sr_index(int sfreq_hz)
{
    return sfreq_hz == 32000 ? 0 : sfreq_hz == 44100 ? 1 : sfreq_hz == 48000 ? 2 : -1;
}

void frame_geometry(int sfreq_hz, int n_ch, int bitrate_kbps, FrameGeom *G)
{
    double avg_slots = (1152.0 / ((double)sfreq_hz / 1000.0)) * ((double)bitrate_kbps / 8.0);
    int whole = (int)avg_slots;
    G->n_ch = n_ch;
    G->bits_per_frame = 8 * whole;
    int sideinfo_len = 32 + (n_ch == 1 ? 136 : 256);
    G->mean_bits = (G->bits_per_frame - sideinfo_len) / 2;
    frame_geom_derive(G);
}

// Emit-side Huffman tables + emission order of short blocks + the constant 32 header bits
// (l3bitstream.c:323-336 with the reference's defaults: no CRC, no padding, not copyrighted, not original, no emphasis;
// sampling_frequency index order 44.1/48/32 and mode 0 = stereo, 3 = mono as in common.c / musicin.c)
void build_bit_tables(int sr, int sfreq_hz, int n_ch, int bitrate_kbps, BitTables *B)
{
    memset(B, 0, sizeof(*B));
    for (int i = 0; i < MP3T_HUFF_FLAT; i++) B->hcode[i] = (MP3T_HCODE[i] & 0xffffffu) | ((unsigned)MP3T_HLEN[i] << 24);
    for (int t = 0; t < 34; t++) {
        B->hoff[t] = MP3T_HUFF[t].off;
        B->ylen[t] = MP3T_HUFF[t].ylen;
        B->linbits[t] = MP3T_HUFF[t].linbits;
    }
    int p = 0;
    for (int sfb = 0; sfb < 13; sfb++)
        for (int w = 0; w < 3; w++)
            for (int line = MP3T_SFB_SHORT[sr][sfb]; line < MP3T_SFB_SHORT[sr][sfb + 1]; line += 2) B->short_e0[p++] = (unsigned short)(line * 3 + w);
    for (int i = 0; i < 23; i++) B->sfb_l[i] = MP3T_SFB_LONG[sr][i];
    static const int rates[15] = {0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320};
    int br = 0;
    for (int i = 1; i < 15; i++) if (rates[i] == bitrate_kbps) br = i;
    const int sf = (sfreq_hz == 44100) ? 0 : (sfreq_hz == 48000) ? 1 : 2;
    const unsigned h = (0xfffu << 20) | (1u << 19) | (1u << 17) | (1u << 16) | ((unsigned)br << 12) | ((unsigned)sf << 10) |
                       ((n_ch == 2 ? 0u : 3u) << 6);
    B->header[0] = (unsigned char)(h >> 24); B->header[1] = (unsigned char)(h >> 16);
    B->header[2] = (unsigned char)(h >> 8); B->header[3] = (unsigned char)h;
    FrameGeom G;
    frame_geometry(sfreq_hz, n_ch, bitrate_kbps, &G);
    B->frame_bytes = G.bits_per_frame / 8;
    B->si_bytes = 4 + (n_ch == 2 ? 32 : 17);
    B->n_ch = n_ch;
}

void build_front_tables(FrontTables *F)
{
    memset(F, 0, sizeof(*F));
    for (int i = 0; i < 512; i++) F->window[i] = MP3T_ANA_WINDOW[i];
    // create_ana_filter, encode.c:331-345: cos rounded to 9 decimals through modf
    for (int sb = 0; sb < 32; sb++) {
        double row[64];
        for (int k = 0; k < 64; k++) {
            double v = 1e9 * cos((double)((2 * sb + 1) * (16 - k) * kRefPi / 64)), ip;
            if (v >= 0) modf(v + 0.5, &ip); else modf(v - 0.5, &ip);
            row[k] = ip * 1e-9;
        }
        for (int j = 0; j < 16; j++) F->am[sb][j] = row[j];           // pairs with ysum[j], encode.c:403
        for (int j = 0; j < 15; j++) F->am[sb][16 + j] = row[33 + j]; // pairs with ysub[j], encode.c:404-406
    }
    static const double c[8] = {-0.6, -0.535, -0.33, -0.185, -0.095, -0.041, -0.0142, -0.0037};  // Table B.9
    for (int k = 0; k < 8; k++) {
        double sq = sqrt(1.0 + c[k] * c[k]);
        F->ca[k] = c[k] / sq;
        F->cs[k] = 1.0 / sq;
    }
    // mdct.c:132-156
    for (int i = 0; i < 36; i++) F->win[0][i] = sin(kRefPi / 36 * (i + 0.5));
    for (int i = 0; i < 18; i++) F->win[1][i] = sin(kRefPi / 36 * (i + 0.5));
    for (int i = 18; i < 24; i++) F->win[1][i] = 1.0;
    for (int i = 24; i < 30; i++) F->win[1][i] = sin(kRefPi / 12 * (i + 0.5 - 18));
    for (int i = 30; i < 36; i++) F->win[1][i] = 0.0;
    for (int i = 0; i < 6; i++) F->win[3][i] = 0.0;
    for (int i = 6; i < 12; i++) F->win[3][i] = sin(kRefPi / 12 * (i + 0.5 - 6));
    for (int i = 12; i < 18; i++) F->win[3][i] = 1.0;
    for (int i = 18; i < 36; i++) F->win[3][i] = sin(kRefPi / 36 * (i + 0.5));
    for (int i = 0; i < 12; i++) F->win[2][i] = sin(kRefPi / 12 * (i + 0.5));
    // mdct.c:158-168
    int N = 12;
    for (int m = 0; m < N / 2; m++)
        for (int k = 0; k < N; k++)
            F->cos_s[m][k] = cos((kRefPi / (2 * N)) * (2 * k + 1 + N / 2) * (2 * m + 1)) / (N / 4);
    N = 36;
    for (int m = 0; m < N / 2; m++)
        for (int k = 0; k < N; k++)
            F->cos_l[m][k] = cos((kRefPi / (2 * N)) * (2 * k + 1 + N / 2) * (2 * m + 1)) / (N / 4);
    for (int m = 0; m < 18; m++)
        for (int j = 0; j < 18; j++) F->dct4_l[m][j] = cos((kRefPi / 72) * (2 * j + 1) * (2 * m + 1)) / 9;
    for (int m = 0; m < 6; m++)
        for (int j = 0; j < 6; j++) F->dct4_s[m][j] = cos((kRefPi / 24) * (2 * j + 1) * (2 * m + 1)) / 3;
}

void build_rate_tables(int sr, RateTables *R)
{
    memset(R, 0, sizeof(*R));
    for (int i = 1; i <= 2048; i++) R->pow_nint_tab[i] = pow((double)i - 0.4054, 4.0 / 3.0);
    for (int i = 0; i < 2048; i++) R->pow43[i] = pow((double)i, 4.0 / 3.0);
    for (int q = -256; q < 256; q++) {
        double step = (q == 0) ? 1.0 : pow(2.0, (double)q * 0.25);
        R->step[q + 256] = step;
        R->ostep[q + 256] = 1.0 / step;
        R->ostep34[q + 256] = (float)pow(1.0 / step, 0.75);
    }
    RateHot &H = R->hot;
    const double ifq = sqrt(2.);
    for (int n = 0; n < 4; n++) {
        H.pre1[n] = pow(ifq, (double)n);
        H.pre2[n] = pow(ifq, 2.0 * (double)n);
    }
    H.ifqstep = sqrt(2.0);
    H.ifqstep2 = H.ifqstep * H.ifqstep;
    H.log2c = log(2.0);
    for (int i = 0; i < 23; i++) H.sfb_l[i] = MP3T_SFB_LONG[sr][i];
    for (int i = 0; i < 14; i++) H.sfb_s[i] = MP3T_SFB_SHORT[sr][i];
    for (int i = 0; i < MP3T_HUFF_FLAT; i++) R->hlen[i] = MP3T_HLEN[i];
    for (int t = 0; t < 34; t++) {
        R->hoff[t] = MP3T_HUFF[t].off;
        R->hxlen[t] = (t >= 32) ? 0 : MP3T_HUFF[t].ylen;   // count1 tables are indexed by p alone
        R->hlinbits[t] = H.hlinbits[t] = MP3T_HUFF[t].linbits;
        R->hlinmax[t] = H.hlinmax[t] = MP3T_HUFF[t].linmax;
    }
    // glut: code length + sign bits of (x, y) in every candidate table of a group, 10-bit fields
    // (count_bit, loop.c:172-225: sum += hlen[x][y] + (x != 0) + (y != 0), escapes add linbits)
    {
        static const int members[8][3] = {{1, 0, 0}, {2, 3, 0}, {5, 6, 0}, {7, 8, 9}, {10, 11, 12}, {13, 15, 0}, {16, 24, -1}, {15, 24, -1}};
        static const int gmax[8] = {1, 2, 3, 5, 7, 14, 15, 15};
        for (int g = 0; g < 8; g++)
            for (int x = 0; x <= gmax[g]; x++)
                for (int y = 0; y < 16; y++) {
                    unsigned v = 0;
                    for (int c = 0; c < 3; c++) {
                        const int t = members[g][c];
                        unsigned f = 0;
                        if (t > 0) {
                            const int xl = MP3T_HUFF[t].xlen, yl = MP3T_HUFF[t].ylen;
                            if (x < xl && y < yl) f = MP3T_HLEN[MP3T_HUFF[t].off + x * yl + y] + (x != 0) + (y != 0);
                        } else if (t < 0) f = (x > 14) + (y > 14);   // number of escaped values of the pair
                        v |= f << (10 * c);
                    }
                    H.glut[16 * GLUT_OFF16(g) + 16 * x + y] = v;
                }
        for (int p = 0; p < 16; p++) {
            const int sg = (p & 1) + ((p >> 1) & 1) + ((p >> 2) & 1) + ((p >> 3) & 1);
            H.c1lut[p] = (unsigned)(MP3T_HLEN[MP3T_HUFF[32].off + p] + sg) | ((unsigned)(MP3T_HLEN[MP3T_HUFF[33].off + p] + sg) << 16);
        }
    }
    for (int s = 0; s < 288; s++) {
        int e = 2 * s, b = 21;
        for (int sfb = 0; sfb < 21; sfb++)
            if (e >= H.sfb_l[sfb] && e < H.sfb_l[sfb + 1]) b = sfb;
        H.band_long[s] = (unsigned char)b;
        int m = s / 3, w = s % 3, line = 2 * m;
        b = 36 + w;
        for (int sfb = 0; sfb < 12; sfb++)
            if (line >= H.sfb_s[sfb] && line < H.sfb_s[sfb + 1]) b = 3 * sfb + w;
        H.band_short[s] = (unsigned char)b;
    }
    {   // runs of band_sums (long blocks): the smallest run limit m with sum_b ceil(width_b / m) <= 32
        int wdt[21], m = 1;
        for (int b = 0; b < 21; b++) wdt[b] = (H.sfb_l[b + 1] >> 1) - (H.sfb_l[b] >> 1);
        for (;; m++) {
            int lanes = 0;
            for (int b = 0; b < 21; b++) lanes += (wdt[b] + m - 1) / m;
            if (lanes <= 32) break;
        }
        int l = 0;
        for (int b = 0; b < 21; b++) {
            const int k = (wdt[b] + m - 1) / m;          // runs of this band (<= 4 for Table B.8), lengths as equal as possible
            int st = H.sfb_l[b] >> 1;
            H.bs_first[b] = (unsigned char)l;
            for (int r = 0; r < k; r++, l++) {
                const int len = wdt[b] / k + (r < wdt[b] % k ? 1 : 0);
                H.bs_start[l] = (unsigned short)st; H.bs_count[l] = (unsigned char)len; H.bs_more[l] = (unsigned char)(r == 0 ? k - 1 : 0);
                st += len;
            }
        }
        for (; l < 32; l++) { H.bs_start[l] = 0; H.bs_count[l] = 0; H.bs_more[l] = 0; }
    }
    // subdivide() for plain long blocks, loop.c:1596-1690, tabulated over big_values
    static const unsigned char subdv[23][2] = {{0,0},{0,0},{0,0},{0,0},{0,0},{0,1},{1,1},{1,1},{1,2},{2,2},{2,3},{2,3},
        {3,4},{3,4},{3,4},{4,5},{4,5},{4,6},{5,6},{5,6},{5,7},{6,7},{6,7}};
    for (int bv = 1; bv <= 288; bv++) {
        int bvr = 2 * bv, n = 0;
        while (H.sfb_l[n] < bvr) n++;
        int r0 = subdv[n][0], index = r0 + 1;
        while (r0 && H.sfb_l[index] > bvr) { r0--; index--; }
        int r1 = subdv[n][1];
        index = r0 + r1 + 2;
        while (r1 && H.sfb_l[index] > bvr) { r1--; index--; }
        H.subdiv[bv][0] = (unsigned char)r0;
        H.subdiv[bv][1] = (unsigned char)r1;
    }
    static const unsigned char pretab[21] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 3, 2};  // Table B.6
    for (int i = 0; i < 21; i++) H.pretab[i] = pretab[i];
}

static double spread_value(double bi, double bj, bool j_ge_i)
{
    // l3psy.c:823-842
    double tempx = j_ge_i ? (bi - bj) * 3.0 : (bi - bj) * 1.5, x, tempy;
    if (tempx >= 0.5 && tempx <= 2.5) {
        double temp = tempx - 0.5;
        x = 8.0 * (temp * temp - 2.0 * temp);
    } else x = 0.0;
    tempx += 0.474;
    tempy = 15.811389 + 7.5 * tempx - 17.5 * sqrt(1.0 + tempx * tempx);
    if (tempy <= -60.0) return 0.0;
    return exp((x + tempy) * kLnToLog10);
}

void build_psy_tables(int sr, PsyTables *P)
{
    memset(P, 0, sizeof(*P));
    const mp3t_part_long &L = MP3T_PART_LONG[sr];
    const mp3t_part_short &S = MP3T_PART_SHORT[sr];
    const mp3t_sfb_map &ML = MP3T_SFBMAP_LONG[sr], &MS = MP3T_SFBMAP_SHORT[sr];
    P->sr_idx = sr; P->n_l = L.n; P->n_s = S.n;
    for (int i = 0; i < 1024; i++) P->hann_l[i] = (float)(0.5 * (1 - cos(2.0 * kRefPi * (i - 0.5) / 1024)));  // l3psy.c:194
    for (int i = 0; i < 256; i++) P->hann_s[i] = (float)(0.5 * (1 - cos(2.0 * kRefPi * (i - 0.5) / 256)));    // l3psy.c:195
    int k2 = 0;
    for (int i = 0; i < L.n; i++) {
        P->numlines_pe[i] = L.lines[i];
        P->minval[i] = L.minval[i]; P->qthr_l[i] = L.qthr[i]; P->norm_l[i] = L.norm[i];
        P->lo_l[i] = (short)k2;
        for (int k = 0; k < L.lines[i]; k++) P->part_l[k2++] = (short)i;
        P->hi_l[i] = (short)k2;
    }
    P->tail_l = k2;  // lines k2..512 keep the zero-initialised partition 0 (l3psy.c:93,805-806)
    for (int i = 0; i < L.n; i++)
        for (int j = 0; j < L.n; j++) P->s3_l[i * 64 + j] = spread_value(L.bval[i], L.bval[j], j >= i);
    for (int i = 0; i < 63; i++)
        for (int j = 0; j < 64; j++) P->s3_lT[j * 64 + i] = P->s3_l[i * 64 + j];
    k2 = 0;
    for (int i = 0; i < S.n; i++) {
        P->numlines_pe[i] = S.lines[i];  // l3psy.c:868 overwrites the long counts read at :796
        P->qthr_s[i] = S.qthr[i]; P->norm_s[i] = S.norm[i];
        P->snr_s_exp[i] = exp((double)S.snr[i] * kLnToLog10);
        P->lo_s[i] = (short)k2;
        for (int k = 0; k < S.lines[i]; k++) P->part_s[k2++] = (short)i;
        P->hi_s[i] = (short)k2;
    }
    for (int i = S.n; i < 42; i++) P->snr_s_exp[i] = exp(0.0 * kLnToLog10);
    P->tail_s = k2;
    P->sparse = (sr == 1);
    // sprdngf1/2 row ranges for 44.1 kHz, l3psy.c:996-1060
    static const unsigned char lo441[63] = {0,0,0,0,0,0,0,0,0,0,0,1,1,2,3,5,6,7,9,10,11,12,14,15,15,16,16,17,18,19,19,20,
        21,22,22,23,24,25,26,27,28,29,30,31,32,33,34,35,36,37,37,38,39,40,41,42,43,44,45,46,47,48,48};
    static const unsigned char hi441[63] = {2,3,4,5,6,7,8,9,10,11,12,14,14,15,15,16,17,19,20,21,22,23,24,25,27,28,28,29,30,31,32,34,
        35,36,36,37,38,39,41,42,43,44,45,46,47,48,49,50,51,52,53,54,55,56,57,58,59,60,61,62,62,62,62};
    for (int b = 0; b < 63; b++) {
        P->spr_lo[b] = P->sparse ? lo441[b] : 0;
        P->spr_hi[b] = P->sparse ? hi441[b] : 62;
    }
    P->spr_wmax = 0; P->pad_spr = 0;
    for (int i = 0; i < 64 * 64; i++) P->s3_band[i] = 1.0;
    for (int b = 0; b < 63; b++) {
        P->spr_wmax = std::max(P->spr_wmax, P->spr_hi[b] - P->spr_lo[b] + 1);
        for (int k = P->spr_lo[b]; k <= P->spr_hi[b]; k++) P->s3_band[(k - P->spr_lo[b]) * 64 + b] = P->s3_lT[k * 64 + b];
    }
    for (int i = 0; i < 21; i++) { P->bu_l[i] = ML.bu[i]; P->bo_l[i] = ML.bo[i]; P->w1_l[i] = ML.w1[i]; P->w2_l[i] = ML.w2[i]; }
    for (int i = 0; i < 12; i++) { P->bu_s[i] = MS.bu[i]; P->bo_s[i] = MS.bo[i]; P->w1_s[i] = MS.w1[i]; P->w2_s[i] = MS.w2[i]; }
    P->n_hist_part = P->part_l[5] + 1;
    P->ch_wmax[0] = P->ch_wmax[1] = 0;
    for (int h = 0; h < 2; h++)
        for (int l = 0; l < 32; l++) {
            const int p = l + 32 * h;
            int lo = p < P->n_l ? P->lo_l[p] : 0, hi = p < P->n_l ? P->hi_l[p] : 0;
            if (p == 63) { lo = P->tail_l; hi = 513; }       // n_l <= 63 for all three sampling rates
            P->ch_lo[h][l] = (short)lo; P->ch_hi[h][l] = (short)hi;
            P->ch_wmax[h] = std::max(P->ch_wmax[h], hi - lo);
        }
}

// ---------------------------------------------------------------------------------------------------
// FFT program builder
// ---------------------------------------------------------------------------------------------------
void build_fft_twiddles(std::vector<FftTwiddle> *tw, std::vector<int> *tw_base)
{
    // subs.c:255-279 / 446-460: `ang` is a float, cos/sin are evaluated in double on it, results are
    // stored as float; -(s+c) and s-c are float operations.  Layout here: per logm >= 4, first the
    // (cn,spcn,smcn) triples for n = 1..m/4-1 (n != m/8), then the (c3n,spc3n,smc3n) triples.
    tw->clear();
    tw_base->assign(11, -1);
    for (int logm = 4; logm <= 10; logm++) {
        int m = 1 << logm, m4 = m / 4, m8 = m / 8;
        (*tw_base)[logm] = (int)tw->size();
        for (int pass = 0; pass < 2; pass++)
            for (int n = 1; n < m4; n++) {
                if (n == m8) continue;
                float ang = (float)((pass ? 3 * n : n) * kTwoPi / m);
                float c = (float)cos((double)ang), s = (float)sin((double)ang);  // double libm on the float angle
                FftTwiddle t;
                t.cn = c; t.spcn = -(s + c); t.smcn = s - c; t.pad = 0.f;
                tw->push_back(t);
            }
    }
}

namespace {
struct Builder {
    std::vector<int> phys;      // logical position -> physical slot
    std::vector<uint8_t> neg;   // logical position currently stored negated
    std::vector<FftOp> ops;
    const std::vector<int> *tw_base;

    void emit(uint8_t type, int a, int b, int c, int d, int tw)
    {
        FftOp o;
        memset(&o, 0, sizeof(o));
        o.type = type; o.tw = (uint16_t)tw;
        int idx[4] = {a, b, c, d};
        uint16_t *dst[4] = {&o.a, &o.b, &o.c, &o.d};
        for (int i = 0; i < 4; i++)
            if (idx[i] >= 0) {
                *dst[i] = (uint16_t)phys[idx[i]];
                if (neg[idx[i]]) o.neg |= 1 << i;
                neg[idx[i]] = 0;  // results are stored with their true sign
            } else *dst[i] = 0xffff;
        ops.push_back(o);
    }
    void flip(int i) { neg[i] ^= 1; }
    void swap_pos(int i, int j) { std::swap(phys[i], phys[j]); std::swap(neg[i], neg[j]); }

    // complex split-radix node on logical positions xr[0..m), xi[0..m)  (srrec, subs.c:185-362)
    void cplx(int xr, int xi, int logm)
    {
        if (logm <= 0) return;
        if (logm == 1) { emit(FFT_BFLY, xr, xr + 1, -1, -1, 0); emit(FFT_BFLY, xi, xi + 1, -1, -1, 0); return; }
        if (logm == 2) {  // subs.c:202-238
            emit(FFT_BFLY, xr, xr + 2, -1, -1, 0); emit(FFT_BFLY, xi, xi + 2, -1, -1, 0);
            emit(FFT_BFLY, xr + 1, xr + 3, -1, -1, 0); emit(FFT_BFLY, xi + 1, xi + 3, -1, -1, 0);
            emit(FFT_BFLY, xr, xr + 1, -1, -1, 0); emit(FFT_BFLY, xi, xi + 1, -1, -1, 0);
            emit(FFT_CROSS, xr + 2, xr + 3, xi + 2, xi + 3, 0);
            return;
        }
        int m = 1 << logm, m2 = m / 2, m4 = m / 4, m8 = m / 8;
        for (int n = 0; n < m2; n++) { emit(FFT_BFLY, xr + n, xr + n + m2, -1, -1, 0); emit(FFT_BFLY, xi + n, xi + n + m2, -1, -1, 0); }
        for (int n = 0; n < m4; n++) emit(FFT_CROSS, xr + m2 + n, xr + m2 + m4 + n, xi + m2 + n, xi + m2 + m4 + n, 0);
        int nel = m4 - 2, k = 0;
        for (int n = 1; n < m4; n++) {
            int a = xr + m2 + n, b = xr + m2 + m4 + n, c = xi + m2 + n, d = xi + m2 + m4 + n;
            if (n == m8) {
                emit(FFT_ROT8A, a, -1, c, -1, 0);
                emit(FFT_ROT8B, b, -1, d, -1, 0);
            } else {
                int base = (*tw_base)[logm];
                emit(FFT_ROT, a, -1, c, -1, base + k);
                emit(FFT_ROT, b, -1, d, -1, base + nel + k);
                k++;
            }
        }
        cplx(xr, xi, logm - 1);
        cplx(xr + m2, xi + m2, logm - 2);
        cplx(xr + 3 * m4, xi + 3 * m4, logm - 2);
    }

    // real split-radix node on logical positions x[0..m)  (rsrec, subs.c:412-523)
    void real(int x, int logm)
    {
        if (logm <= 0) return;
        if (logm == 1) { emit(FFT_BFLY, x, x + 1, -1, -1, 0); return; }
        int m = 1 << logm, m2 = m / 2, m4 = m / 4, m8 = m / 8;
        for (int n = 0; n < m2; n++) emit(FFT_BFLY, x + n, x + n + m2, -1, -1, 0);
        for (int n = 0; n < m4; n++) flip(x + m2 + m4 + n);
        int k = 0;
        for (int n = 1; n < m4; n++) {
            int a = x + m2 + n, c = x + m2 + m4 + n;
            if (n == m8) emit(FFT_ROT8A, a, -1, c, -1, 0);
            else { emit(FFT_ROT, a, -1, c, -1, (*tw_base)[logm] + k); k++; }
        }
        real(x, logm - 1);
        cplx(x + m2, x + 3 * m4, logm - 2);
        {   // step 5, subs.c:501-521: new p = -old q, new q = -old p ; then new p = -old q, new q = old p
            int p = x + m2 + m4, q = x + m - 1;
            for (int n = 0; n < m8; n++) { swap_pos(p, q); flip(p); flip(q); p++; q--; }
            p = x + m2 + 1; q = x + m - 2;
            for (int n = 0; n < m8; n++) { swap_pos(p, q); flip(p); p += 2; q -= 2; }
        }
        if (logm == 2) flip(x + 3);
    }
};
}  // namespace

namespace {
// banks of the (up to four) operands of an op
inline void op_banks(const FftOp &o, int bank[4])
{
    const uint16_t s4[4] = {o.a, o.b, o.c, o.d};
    for (int j = 0; j < 4; j++) bank[j] = s4[j] == 0xffff ? -1 : (int)(FFT_SKEW((unsigned)s4[j]) & 31);
}

// Conflict-free rows that are not full are merged (first-fit decreasing, capacity 32 ops): a row merged from k
// matchings has at most k-way bank conflicts, i.e. it costs the same shared-memory wavefronts as its k parts did, but
// only one row's worth of instructions.
void merge_rows(std::vector<std::vector<int>> *rows)
{
    std::vector<std::vector<int>> in = *rows, out;
    std::stable_sort(in.begin(), in.end(), [](const std::vector<int> &x, const std::vector<int> &y) { return x.size() > y.size(); });
    for (std::vector<int> &r : in) {
        size_t k = 0;
        while (k < out.size() && out[k].size() + r.size() > 32) k++;
        if (k == out.size()) out.emplace_back();
        out[k].insert(out[k].end(), r.begin(), r.end());
    }
    *rows = out;
}

// Split the ops `ids` (one dependency level, one operand class) into rows of <= 32 ops without bank conflicts.
// Two-operand classes: an op is an edge between the bank of its first and the bank of its second operand, a
// conflict-free row is a matching, and a bipartite multigraph of maximum degree D splits into exactly D matchings
// (Koenig) — found with the alternating-path recolouring below: the minimum number of rows.  Four-operand crosses:
// first-fit over many random orders, best result kept.
void pack_rows(const std::vector<FftOp> &ops, const std::vector<int> &ids, bool four, std::vector<std::vector<int>> *rows)
{
    rows->clear();
    if (ids.empty()) return;
    if (!four) {
        const int ne = (int)ids.size();
        std::vector<int> eu(ne), ev(ne), col(ne, -1);
        int degu[32] = {0}, degv[32] = {0}, D = 0;
        for (int e = 0; e < ne; e++) {
            int b[4]; op_banks(ops[ids[e]], b);
            eu[e] = b[0]; ev[e] = b[1] >= 0 ? b[1] : b[2];
            D = std::max(D, std::max(++degu[eu[e]], ++degv[ev[e]]));
        }
        std::vector<std::vector<int>> atu(D, std::vector<int>(32, -1)), atv(D, std::vector<int>(32, -1));   // [colour][bank] -> edge
        for (int e = 0; e < ne; e++) {
            const int u = eu[e], v = ev[e];
            int c1 = 0, c2 = 0;
            while (atu[c1][u] >= 0) c1++;
            while (atv[c2][v] >= 0) c2++;
            if (atv[c1][v] >= 0) {
                // c1 is taken at v: swap c1 <-> c2 along the alternating path that starts at v (it cannot reach u)
                std::vector<int> path;
                int cur = v, ca = c1, cb = c2;
                bool right = true;
                for (;;) {
                    const int e1 = right ? atv[ca][cur] : atu[ca][cur];
                    if (e1 < 0) break;
                    path.push_back(e1);
                    cur = right ? eu[e1] : ev[e1];
                    right = !right;
                    std::swap(ca, cb);
                }
                for (int e1 : path) { atu[col[e1]][eu[e1]] = -1; atv[col[e1]][ev[e1]] = -1; }
                for (int e1 : path) { col[e1] = (col[e1] == c1) ? c2 : c1; atu[col[e1]][eu[e1]] = e1; atv[col[e1]][ev[e1]] = e1; }
            }
            col[e] = c1; atu[c1][u] = e; atv[c1][v] = e;
        }
        rows->assign(D, std::vector<int>());
        for (int e = 0; e < ne; e++) (*rows)[col[e]].push_back(ids[e]);
        merge_rows(rows);
        return;
    }
    std::vector<int> perm = ids;
    std::vector<std::vector<int>> best;
    uint64_t rng = 0x9e3779b97f4a7c15ull;
    for (int trial = 0; trial < 200; trial++) {
        if (trial)
            for (size_t k = perm.size() - 1; k > 0; k--) {
                rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                std::swap(perm[k], perm[(size_t)((rng >> 33) % (k + 1))]);
            }
        std::vector<std::vector<int>> cur;
        std::vector<std::array<uint32_t, 4>> used;
        for (int id : perm) {
            int b[4]; op_banks(ops[id], b);
            size_t r = 0;
            for (; r < cur.size(); r++) {
                if (cur[r].size() >= 32) continue;
                bool ok = true;
                for (int j = 0; j < 4; j++) if (b[j] >= 0 && (used[r][j] >> b[j] & 1)) ok = false;
                if (ok) break;
            }
            if (r == cur.size()) { cur.emplace_back(); used.push_back({0, 0, 0, 0}); }
            cur[r].push_back(id);
            for (int j = 0; j < 4; j++) if (b[j] >= 0) used[r][j] |= 1u << b[j];
        }
        if (best.empty() || cur.size() < best.size()) best = cur;
        if (best.size() * 32 < ids.size() + 32) break;   // cannot do better
    }
    *rows = best;
    merge_rows(rows);
}
}  // namespace

void build_fft_program(int logm, const std::vector<int> &tw_base, FftProgram *P)
{
    const int n = 1 << logm;
    Builder B;
    B.phys.resize(n); B.neg.assign(n, 0); B.tw_base = &tw_base;
    for (int i = 0; i < n; i++) B.phys[i] = i;
    B.real(0, logm);
    // BR_permute, subs.c:136-177 (Evans' algorithm == full bit reversal for even logm)
    for (int i = 0; i < n; i++) {
        int r = 0;
        for (int b = 0; b < logm; b++) if (i & (1 << b)) r |= 1 << (logm - 1 - b);
        if (r > i) B.swap_pos(i, r);
    }
    // levelise: every op reads and writes all of its slots
    std::vector<int> last(n, 0), level(B.ops.size());
    int n_levels = 0;
    for (size_t i = 0; i < B.ops.size(); i++) {
        const FftOp &o = B.ops[i];
        int l = 0;
        const uint16_t s[4] = {o.a, o.b, o.c, o.d};
        for (int j = 0; j < 4; j++) if (s[j] != 0xffff) l = std::max(l, last[s[j]]);
        l += 1;
        for (int j = 0; j < 4; j++) if (s[j] != 0xffff) last[s[j]] = l;
        level[i] = l;
        n_levels = std::max(n_levels, l);
    }
    std::vector<int> order(B.ops.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
        if (level[x] != level[y]) return level[x] < level[y];
        return B.ops[x].type < B.ops[y].type;
    });
    P->n = n; P->logm = logm;
    P->ops.clear(); P->level_start.assign(FFT_CLASSES * n_levels + 1, 0);
    P->words.clear(); P->seg_word.assign(FFT_CLASSES * n_levels + 1, 0);
    // Pack every (level, class) group into rows of 32 ops (one op per lane) such that no two ops of a row touch the same
    // shared-memory bank with the same operand: each of the row's loads and stores is then a single wavefront.
    {
        FftOp nop; memset(&nop, 0, sizeof(nop)); nop.a = nop.b = nop.c = nop.d = 0xffff; nop.type = FFT_NOP;
        size_t i = 0;
        for (int l = 1; l <= n_levels; l++) {
            std::vector<int> grp[FFT_CLASSES];
            for (; i < order.size() && level[order[i]] == l; i++) grp[fft_class(B.ops[order[i]])].push_back(order[i]);
            for (int c = 0; c < FFT_CLASSES; c++) {
                std::vector<std::vector<int>> rows;
                pack_rows(B.ops, grp[c], c == 1, &rows);
                for (const std::vector<int> &row : rows) {
                    uint32_t used[4] = {0, 0, 0, 0};
                    for (int id : row) {
                        P->ops.push_back(B.ops[id]);
                        int b[4]; op_banks(B.ops[id], b);
                        if (c == 3 && b[1] < 0) b[1] = b[2];           // MISC ops carry their second operand in position 1
                        if (c == 2) b[1] = b[2];
                        for (int j = 0; j < 4; j++) if (b[j] >= 0 && (c == 1 || j < 2)) used[j] |= 1u << b[j];
                    }
                    // padding lanes: dummy words in banks the row's real ops leave free (operand j of the lane: pad byte j)
                    for (size_t k = row.size(); k < 32; k++) {
                        FftOp pd = nop;
                        for (int j = 0; j < 4; j++) {
                            int f = 0;
                            while (used[j] >> f & 1) f++;
                            used[j] |= 1u << f;
                            pd.pad |= (uint32_t)f << (8 * j);
                        }
                        P->ops.push_back(pd);
                    }
                }
                P->level_start[FFT_CLASSES * (l - 1) + c + 1] = (int)P->ops.size();   // end of (level, class) segment
            }
        }
    }
    // device encoding
    std::vector<FftTwiddle> tw_all; std::vector<int> tw_base_chk;
    build_fft_twiddles(&tw_all, &tw_base_chk);
    for (int sgm = 0; sgm < FFT_CLASSES * n_levels; sgm++) {
        const int c = sgm % FFT_CLASSES;
        for (int k = P->level_start[sgm]; k < P->level_start[sgm + 1]; k++) {
            const FftOp &o = P->ops[k];
            auto off = [&](uint16_t slot, int j) -> uint32_t {      // byte offset of an operand; padding -> the lane's dummy word j
                // dummy word in the free bank f picked above: the dummy area starts at word n + n / 32 right behind the data
                const unsigned dummy0 = (unsigned)FFT_SKEW((unsigned)n);
                return 4u * (o.type == FFT_NOP ? dummy0 + 32u * j + ((((o.pad >> (8 * j)) & 31u) - dummy0) & 31u) : (unsigned)FFT_SKEW((unsigned)slot));
            };
            if (c == 0) P->words.push_back(off(o.a, 0) | (off(o.b, 1) << 16));
            else if (c == 1) { P->words.push_back(off(o.a, 0) | (off(o.b, 1) << 16)); P->words.push_back(off(o.c, 2) | (off(o.d, 3) << 16)); }
            else if (c == 2) {
                // the twiddle triple travels in the op itself (coalesced with it) instead of being gathered from a table
                FftTwiddle t = {0.f, 0.f, 0.f, 0.f};
                if (o.type != FFT_NOP) t = tw_all[o.tw];
                uint32_t tb[3];
                memcpy(&tb[0], &t.cn, 4); memcpy(&tb[1], &t.spcn, 4); memcpy(&tb[2], &t.smcn, 4);
                P->words.push_back(off(o.a, 0) | (off(o.c, 1) << 16) | ((o.neg & 4) ? 0x80000000u : 0u));
                P->words.push_back(tb[0]); P->words.push_back(tb[1]); P->words.push_back(tb[2]);
            } else {
                const uint16_t second = (o.type == FFT_BFLY) ? o.b : o.c;
                const unsigned neg2 = (o.type == FFT_BFLY) ? ((o.neg >> 1) & 1) : ((o.neg >> 2) & 1);
                P->words.push_back(off(o.a, 0) | (off(second, 1) << 16));
                P->words.push_back((uint32_t)o.type | ((uint32_t)((o.neg & 1) | (neg2 << 1)) << 8));
            }
        }
        P->seg_word[sgm + 1] = (int)P->words.size();
    }
    P->words.resize(P->words.size() + FFT_WORDS_PAD, 0u);   // the executor fetches op words one trip ahead, past the last segment
    P->out_slot.resize(n); P->out_neg.resize(n);
    for (int i = 0; i < n; i++) { P->out_slot[i] = (uint16_t)B.phys[i]; P->out_neg[i] = B.neg[i]; }
}

// ---------------------------------------------------------------------------------------------------
// FFT in registers: block kinds, lane / slot assignment, twiddle tables, output maps (see tables.h)
// ---------------------------------------------------------------------------------------------------
namespace {
struct RegsBlock { int kind, partner, is_xi; };
struct RegsWalker {
    RegsBlock blk[32];
    // real node on elements [base, base + 2^logm), rsrec subs.c:412-523: its lower half is a real node, its upper half a
    // complex node of a quarter of the length (real array at base + m/2, imaginary array at base + 3m/4)
    void rs(int base, int logm)
    {
        if (logm == 6) { blk[base / 32] = {FFTR_C, -1, 0}; blk[base / 32 + 1] = {FFTR_D, -1, 0}; return; }
        const int m = 1 << logm;
        rs(base, logm - 1);
        sr(base + m / 2, base + 3 * m / 4, logm - 2);
    }
    // complex node, srrec subs.c:185-362
    void sr(int xr, int xi, int logm)
    {
        if (logm == 5) { blk[xr / 32] = {FFTR_A, xi / 32, 0}; blk[xi / 32] = {FFTR_A, xr / 32, 1}; return; }
        if (logm == 6) {
            sr(xr, xi, 5);
            blk[xr / 32 + 1] = {FFTR_B, xi / 32 + 1, 0}; blk[xi / 32 + 1] = {FFTR_B, xr / 32 + 1, 1};
            return;
        }
        const int m = 1 << logm;
        sr(xr, xi, logm - 1);
        sr(xr + m / 2, xi + m / 2, logm - 2);
        sr(xr + 3 * m / 4, xi + 3 * m / 4, logm - 2);
    }
};
}  // namespace

void build_fft_regs_plan(FftRegsPlan *P)
{
    memset(&P->c, 0, sizeof(P->c));
    std::vector<FftTwiddle> tw; std::vector<int> base;
    build_fft_twiddles(&tw, &base);
    auto triple = [&](int L, int set, int n) -> FftTwC {
        const int m4 = 1 << (L - 2), m8 = m4 / 2, nel = m4 - 2;
        FftTwC t = {0.f, 0.f, 0.f};
        if (n == 0 || n == m8) return t;
        const FftTwiddle &w = tw[base[L] + set * nel + (n < m8 ? n - 1 : n - 2)];
        t.cn = w.cn; t.spcn = w.spcn; t.smcn = w.smcn;
        return t;
    };
    for (int set = 0; set < 2; set++) {
        for (int n = 0; n < 4; n++) P->c.small.t4[set][n] = triple(4, set, n);
        for (int n = 0; n < 8; n++) P->c.small.t5[set][n] = triple(5, set, n);
        for (int n = 0; n < 16; n++) P->c.small.t6[set][n] = triple(6, set, n);
    }
    P->twA.assign((size_t)FFTR_TWA_ENTRIES * 4, 0.f);
    for (int L = 7; L <= 10; L++)
        for (int set = 0; set < 2; set++)
            for (int n = 0; n < (1 << (L - 2)); n++) {
                const FftTwC t = triple(L, set, n);
                float *d = &P->twA[(size_t)(FFTR_TWA_OFF(L, set) + n) * 4];
                d[0] = t.cn; d[1] = t.spcn; d[2] = t.smcn;
            }
    // block kinds; lanes of pass 0 = the kind-a blocks, lanes of pass 1 = kinds b (16), c (4), d (4); slot = 32 * pass + lane
    RegsWalker WL, WS;
    WL.rs(0, 10); WS.rs(0, 8);
    int slot_l[32], slot_s[3][8];
    int next[2] = {0, 0};
    for (int kind = 0; kind < 4; kind++) {
        const int pass = kind == FFTR_A ? 0 : 1;
        for (int b = 0; b < 32; b++) if (WL.blk[b].kind == kind) slot_l[b] = 32 * pass + next[pass]++;
        for (int t = 0; t < 3; t++)
            for (int b = 0; b < 8; b++) if (WS.blk[b].kind == kind) slot_s[t][b] = 32 * pass + next[pass]++;
    }
    // fft_regs.h hard-codes the lane ranges of the two passes: 32 lanes of kind a; kinds b / c / d in lanes 0..15 / 16..19 / 20..23
    {
        int cnt[4] = {0, 0, 0, 0};
        for (int b = 0; b < 32; b++) cnt[WL.blk[b].kind]++;
        for (int b = 0; b < 8; b++) cnt[WS.blk[b].kind] += 3;
        if (cnt[FFTR_A] != 32 || cnt[FFTR_B] != 16 || cnt[FFTR_C] != 4 || cnt[FFTR_D] != 4 || next[0] != 32 || next[1] != 24) {
            fprintf(stderr, "mp3gpu: FFT block plan does not match the kernel's lane layout\n");
            abort();
        }
    }
    for (int b = 0; b < 32; b++) {
        P->kind_long[b] = (uint8_t)WL.blk[b].kind;
        P->c.long_word[b] = (uint16_t)(FFTR_SLOT_WORDS * slot_l[b]);
        if (WL.blk[b].partner >= 0) {
            const int s = slot_l[b], ps = slot_l[WL.blk[b].partner];
            P->c.partner[s / 32][s % 32] = (uint8_t)(ps % 32);
            P->c.is_xi[s / 32][s % 32] = (uint8_t)WL.blk[b].is_xi;
        }
    }
    for (int b = 0; b < 8; b++) P->kind_short[b] = (uint8_t)WS.blk[b].kind;
    for (int t = 0; t < 3; t++)
        for (int b = 0; b < 8; b++) {
            P->c.short_word[t][b] = (uint16_t)(FFTR_SLOT_WORDS * slot_s[t][b]);
            if (WS.blk[b].partner >= 0) {
                const int s = slot_s[t][b], ps = slot_s[t][WS.blk[b].partner];
                P->c.partner[s / 32][s % 32] = (uint8_t)(ps % 32);
                P->c.is_xi[s / 32][s % 32] = (uint8_t)WS.blk[b].is_xi;
            }
        }
    // output maps: every op works in place at the reference's array positions; the permutations of rsrec step 5 and
    // BR_permute and the sign changes that follow the last arithmetic on a slot are exactly what build_fft_program tracks
    FftProgram P10, P8;
    build_fft_program(10, base, &P10);
    build_fft_program(8, base, &P8);
    auto word_l = [&](int j) { return (uint32_t)(P->c.long_word[j >> 5] + (j & 31)); };
    P->out_long.assign(513, 0u);
    for (int i = 0; i <= 512; i++) {
        auto one = [&](int k) { return word_l(P10.out_slot[k]) | (P10.out_neg[k] ? 0x8000u : 0u); };
        P->out_long[i] = one(i) | ((i > 0 ? one(1024 - i) : 0u) << 16);
    }
    P->out_short.assign(3 * 132, 0u);
    for (int t = 0; t < 3; t++)
        for (int i = 0; i <= 128; i++) {
            auto one = [&](int k) { const int j = P8.out_slot[k]; return (uint32_t)(P->c.short_word[t][j >> 5] + (j & 31)) | (P8.out_neg[k] ? 0x8000u : 0u); };
            P->out_short[t * 132 + i] = one(i) | ((i > 0 ? one(256 - i) : 0u) << 16);
        }
}

void run_fft_program_host(const FftProgram &P, const std::vector<FftTwiddle> &tw, float *x)
{
    const double SQ = 0.707106781186547524401;  // subs.c:26
    for (size_t i = 0; i < P.ops.size(); i++) {
        const FftOp &o = P.ops[i];
        float a = 0, b = 0, c = 0, d = 0, t1, t2;
        if (o.type == FFT_NOP) continue;
        if (o.a != 0xffff) a = (o.neg & 1) ? -x[o.a] : x[o.a];
        if (o.b != 0xffff) b = (o.neg & 2) ? -x[o.b] : x[o.b];
        if (o.c != 0xffff) c = (o.neg & 4) ? -x[o.c] : x[o.c];
        if (o.d != 0xffff) d = (o.neg & 8) ? -x[o.d] : x[o.d];
        switch (o.type) {
        case FFT_BFLY: t1 = a + b; b = a - b; a = t1; x[o.a] = a; x[o.b] = b; break;
        case FFT_CROSS: t1 = a + d; t2 = c + b; c = c - b; b = a - d; a = t1; d = t2;
            x[o.a] = a; x[o.b] = b; x[o.c] = c; x[o.d] = d; break;
        case FFT_ROT: {
            const FftTwiddle &w = tw[o.tw];
            t2 = w.cn * (a + c); t1 = w.spcn * a + t2; a = w.smcn * c + t2; c = t1;
            x[o.a] = a; x[o.c] = c; break; }
        case FFT_ROT8A: t1 = (float)(SQ * (a + c)); c = (float)(SQ * (c - a)); a = t1; x[o.a] = a; x[o.c] = c; break;
        case FFT_ROT8B: t2 = (float)(SQ * (c - a)); c = (float)(-SQ * (a + c)); a = t2; x[o.a] = a; x[o.c] = c; break;
        }
    }
}

}  // namespace mp3gpu
