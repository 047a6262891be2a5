/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See l3oracle.h.
 *
 * Sequential CPU restatement of the reference hot path.  Every function cites the reference
 * file:line it follows (paths relative to /root/reference/src).  Arithmetic types follow the
 * reference exactly (FLOAT == float, common.h:183-187): FFT, energies, phases, cb/ecb/nb are
 * float; everything else double.  Build with -ffp-contract=off (oracle/Makefile).
 */
#include "l3oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../mp3-enc-bsd_b200/csrc/iso_tables.h"

#define REF_PI 3.14159265358979      /* common.h:199 (not M_PI) */
#define LN_TO_LOG10 0.2302585093     /* common.h:204 */
#define TWOPI 6.28318530717958647692 /* subs.c:25 */
#define SQHALF 0.707106781186547524401

/* ------------------------------------------------------------------------------------------ */
/* constant tables, computed with the same libm expressions as the reference                  */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int ready;
    double ana_m[32][64];                       /* encode.c:331-345 */
    double win[4][36], cos_l[18][36], cos_s[6][12], ca[8], cs[8]; /* mdct.c:38-44,132-168 */
    float hann_l[1024], hann_s[256];            /* l3psy.c:194-195 */
    float *tw[11];                              /* subs.c:255-279: 6 arrays of m/4-2 per logm>=4 */
    int brev10[1024], brev8[256];
    double pow_nint_tab[4096];                  /* pow_nint.c:13-19 */
    double pow43[2048];                         /* loop.c:1017-1021 (extended: ix<=2047) */
} common_tables;

typedef struct {
    int ready, sr_idx;
    int n_l, n_s;
    int numlines_pe[63];                        /* l3psy.c:796 then :868 (short overwrites long) */
    int part_l[513], part_s[129];
    double minval[63], qthr_l[63], norm_l[63], s3_l[63][63];
    double qthr_s[63], norm_s[63], snr_s[63];
    int bu_l[21], bo_l[21], bu_s[12], bo_s[12];
    double w1_l[21], w2_l[21], w1_s[12], w2_s[12];
} psy_tables;

static common_tables CT;
static psy_tables PT[3];

static void build_fft_twiddles(void)
{
    /* subs.c:255-279 (srrec) and :446-460 (rsrec): float ang, double cos/sin, float results */
    int logm;
    for (logm = 4; logm <= 10; logm++) {
        int m = 1 << logm, m4 = m / 4, m8 = m / 8, nel = m4 - 2, n, k = 0;
        float *t = (float *)calloc((size_t)(6 * nel), sizeof(float));
        for (n = 1; n < m4; n++) {
            float ang, c, s;
            if (n == m8) continue;
            ang = (float)(n * TWOPI / m);
            c = (float)cos(ang); s = (float)sin(ang);
            t[k] = c; t[nel + k] = -(s + c); t[2 * nel + k] = s - c;
            ang = (float)(3 * n * TWOPI / m);
            c = (float)cos(ang); s = (float)sin(ang);
            t[3 * nel + k] = c; t[4 * nel + k] = -(s + c); t[5 * nel + k] = s - c;
            k++;
        }
        CT.tw[logm] = t;
    }
}

static void init_common(void)
{
    int i, k, m, N;
    static const double c_alias[8] = {-0.6, -0.535, -0.33, -0.185, -0.095, -0.041, -0.0142, -0.0037};
    if (CT.ready) return;
    for (i = 0; i < 32; i++)
        for (k = 0; k < 64; k++) {
            double v = 1e9 * cos((double)((2 * i + 1) * (16 - k) * REF_PI / 64));
            double ip;
            if (v >= 0) modf(v + 0.5, &ip); else modf(v - 0.5, &ip);
            CT.ana_m[i][k] = ip * 1e-9;
        }
    for (k = 0; k < 8; k++) {
        double sq = sqrt(1.0 + c_alias[k] * c_alias[k]);
        CT.ca[k] = c_alias[k] / sq;
        CT.cs[k] = 1.0 / sq;
    }
    for (i = 0; i < 36; i++) CT.win[0][i] = sin(REF_PI / 36 * (i + 0.5));
    for (i = 0; i < 18; i++) CT.win[1][i] = sin(REF_PI / 36 * (i + 0.5));
    for (i = 18; i < 24; i++) CT.win[1][i] = 1.0;
    for (i = 24; i < 30; i++) CT.win[1][i] = sin(REF_PI / 12 * (i + 0.5 - 18));
    for (i = 30; i < 36; i++) CT.win[1][i] = 0.0;
    for (i = 0; i < 6; i++) CT.win[3][i] = 0.0;
    for (i = 6; i < 12; i++) CT.win[3][i] = sin(REF_PI / 12 * (i + 0.5 - 6));
    for (i = 12; i < 18; i++) CT.win[3][i] = 1.0;
    for (i = 18; i < 36; i++) CT.win[3][i] = sin(REF_PI / 36 * (i + 0.5));
    for (i = 0; i < 12; i++) CT.win[2][i] = sin(REF_PI / 12 * (i + 0.5));
    for (i = 12; i < 36; i++) CT.win[2][i] = 0.0;
    N = 12;
    for (m = 0; m < N / 2; m++)
        for (k = 0; k < N; k++)
            CT.cos_s[m][k] = cos((REF_PI / (2 * N)) * (2 * k + 1 + N / 2) * (2 * m + 1)) / (N / 4);
    N = 36;
    for (m = 0; m < N / 2; m++)
        for (k = 0; k < N; k++)
            CT.cos_l[m][k] = cos((REF_PI / (2 * N)) * (2 * k + 1 + N / 2) * (2 * m + 1)) / (N / 4);
    for (i = 0; i < 1024; i++) CT.hann_l[i] = (float)(0.5 * (1 - cos(2.0 * REF_PI * (i - 0.5) / 1024)));
    for (i = 0; i < 256; i++) CT.hann_s[i] = (float)(0.5 * (1 - cos(2.0 * REF_PI * (i - 0.5) / 256)));
    build_fft_twiddles();
    for (i = 0; i < 1024; i++) { int r = 0, b; for (b = 0; b < 10; b++) if (i & (1 << b)) r |= 1 << (9 - b); CT.brev10[i] = r; }
    for (i = 0; i < 256; i++) { int r = 0, b; for (b = 0; b < 8; b++) if (i & (1 << b)) r |= 1 << (7 - b); CT.brev8[i] = r; }
    for (i = 1; i < 4096; i++) CT.pow_nint_tab[i] = pow((double)i - 0.4054, 4.0 / 3.0);
    for (i = 0; i < 2048; i++) CT.pow43[i] = pow((double)i, 4.0 / 3.0);
    CT.ready = 1;
}

/* spreading function value, l3psy.c:823-842 */
static double spread_val(double bi, double bj, int j_ge_i)
{
    double tempx, x, tempy, temp;
    tempx = j_ge_i ? (bi - bj) * 3.0 : (bi - bj) * 1.5;
    if (tempx >= 0.5 && tempx <= 2.5) { temp = tempx - 0.5; x = 8.0 * (temp * temp - 2.0 * temp); }
    else x = 0.0;
    tempx += 0.474;
    tempy = 15.811389 + 7.5 * tempx - 17.5 * sqrt(1.0 + tempx * tempx);
    if (tempy <= -60.0) return 0.0;
    return exp((x + tempy) * LN_TO_LOG10);
}

/* L3para_read restated on the structured tables, l3psy.c:770-994 */
static void init_psy(int sr)
{
    psy_tables *p = &PT[sr];
    const mp3t_part_long *L = &MP3T_PART_LONG[sr];
    const mp3t_part_short *S = &MP3T_PART_SHORT[sr];
    const mp3t_sfb_map *ML = &MP3T_SFBMAP_LONG[sr], *MS = &MP3T_SFBMAP_SHORT[sr];
    int i, j, k, k2;
    if (p->ready) return;
    memset(p, 0, sizeof(*p));
    p->sr_idx = sr; p->n_l = L->n; p->n_s = S->n;
    for (i = 0, k2 = 0; i < L->n; i++) {
        p->numlines_pe[i] = L->lines[i];
        p->minval[i] = L->minval[i]; p->qthr_l[i] = L->qthr[i]; p->norm_l[i] = L->norm[i];
        for (k = 0; k < L->lines[i]; k++) p->part_l[k2++] = i;
    }
    for (i = 0; i < L->n; i++)
        for (j = 0; j < L->n; j++)
            p->s3_l[i][j] = spread_val(L->bval[i], L->bval[j], j >= i);
    for (i = 0, k2 = 0; i < S->n; i++) {
        p->numlines_pe[i] = S->lines[i];      /* the quirk: short counts overwrite long ones */
        p->qthr_s[i] = S->qthr[i]; p->norm_s[i] = S->norm[i]; p->snr_s[i] = S->snr[i];
        for (k = 0; k < S->lines[i]; k++) p->part_s[k2++] = i;
    }
    for (i = 0; i < 21; i++) { p->bu_l[i] = ML->bu[i]; p->bo_l[i] = ML->bo[i]; p->w1_l[i] = ML->w1[i]; p->w2_l[i] = ML->w2[i]; }
    for (i = 0; i < 12; i++) { p->bu_s[i] = MS->bu[i]; p->bo_s[i] = MS->bo[i]; p->w1_s[i] = MS->w1[i]; p->w2_s[i] = MS->w2[i]; }
    p->ready = 1;
}

/* ------------------------------------------------------------------------------------------ */
/* polyphase analysis filterbank                                                              */
/* ------------------------------------------------------------------------------------------ */
/* hist[j] = sample (t_newest - j)/32768, i.e. the ring of encode.c:306-314 read in z order */
static void polyphase_slot(double hist[512], const short *pcm32, double s[32])
{
    double z[512], y[64], ysum[16], ysub[16];
    int i, j;
    memmove(hist + 32, hist, 480 * sizeof(double));
    for (i = 0; i < 32; i++) hist[31 - i] = (double)pcm32[i] / 32768;          /* encode.c:306-307 */
    for (i = 0; i < 512; i++) z[i] = hist[i] * MP3T_ANA_WINDOW[i];             /* encode.c:310-311 */
    for (i = 0; i < 64; i++)                                                   /* encode.c:392-396 */
        y[i] = z[i] + z[i + 64] + z[i + 128] + z[i + 192] + z[i + 256] + z[i + 320] + z[i + 384] + z[i + 448];
    for (i = 0; i < 16; i++) ysum[i] = y[i] + y[32 - i];                       /* encode.c:397 */
    for (i = 0; i < 15; i++) ysub[i] = y[33 + i] - y[63 - i];                  /* encode.c:398 */
    for (i = 0; i < 32; i++) {                                                 /* encode.c:399-408 */
        double si = y[16];
        for (j = 0; j < 16; j++) si += CT.ana_m[i][j] * ysum[j];
        for (j = 0; j < 15; j++) si += CT.ana_m[i][33 + j] * ysub[j];
        s[i] = si;
    }
}

void l3o_polyphase(const short *pcm, long n_slots, double *sb)
{
    double hist[512];
    long n;
    init_common();
    memset(hist, 0, sizeof(hist));
    for (n = 0; n < n_slots; n++) polyphase_slot(hist, pcm + 32 * n, sb + 32 * n);
}

/* ------------------------------------------------------------------------------------------ */
/* MDCT + alias reduction                                                                     */
/* ------------------------------------------------------------------------------------------ */
/* one band: mdct.c:105-198 (type 0 uses the same dot product; the reference's hand-unrolled
 * form mdct.c:199-509 is algebraically identical, summation order differs by <=4e-15) */
static void mdct_band(const double in[36], double *out, int bt)
{
    int k, l, m;
    if (bt == 2) {
        for (l = 0; l < 3; l++)
            for (m = 0; m < 6; m++) {
                double sum = 0.0;
                for (k = 0; k < 12; k++) sum += CT.win[2][k] * in[k + 6 * l + 6] * CT.cos_s[m][k];
                out[3 * m + l] = sum;
            }
    } else {
        double fin[36];
        for (k = 0; k < 36; k++) fin[k] = CT.win[bt][k] * in[k];
        for (m = 0; m < 18; m++) {
            double sum = 0.0;
            for (k = 0; k < 36; k++) sum += fin[k] * CT.cos_l[m][k];
            out[m] = sum;
        }
    }
}

/* prev/cur: sign-fixed subband samples [18][32]; mdct.c:63-91 */
static void mdct_granule_fixed(const double *prev, const double *cur, int bt, double *xr)
{
    int band, k;
    double in[36];
    for (band = 0; band < 32; band++) {
        for (k = 0; k < 18; k++) { in[k] = prev[k * 32 + band]; in[k + 18] = cur[k * 32 + band]; }
        mdct_band(in, xr + band * 18, bt);
    }
    if (bt != 2)
        for (band = 0; band < 31; band++)
            for (k = 0; k < 8; k++) {
                double a = xr[band * 18 + 17 - k], b = xr[(band + 1) * 18 + k];
                double bu = a * CT.cs[k] + b * CT.ca[k];
                double bd = b * CT.cs[k] - a * CT.ca[k];
                xr[band * 18 + 17 - k] = bu;
                xr[(band + 1) * 18 + k] = bd;
            }
}

static void sign_fix(const double *raw, double *fixed) /* mdct.c:57-60 */
{
    int k, band;
    for (k = 0; k < 18; k++)
        for (band = 0; band < 32; band++)
            fixed[k * 32 + band] = ((band & 1) && (k & 1)) ? raw[k * 32 + band] * -1.0 : raw[k * 32 + band];
}

void l3o_mdct_granule(const double *prev, const double *cur, int block_type, double *xr)
{
    double p[576], c[576];
    init_common();
    sign_fix(prev, p);
    sign_fix(cur, c);
    mdct_granule_fixed(p, c, block_type, xr);
}

/* ------------------------------------------------------------------------------------------ */
/* FP32 split-radix real FFT (Malvar), subs.c:185-534, restated without statics               */
/* ------------------------------------------------------------------------------------------ */
static void cplx_sr(float *xr, float *xi, int logm) /* srrec, subs.c:185-362 */
{
    int m, m2, m4, m8, n;
    float t1, t2;
    if (logm <= 0) return;
    if (logm == 1) {
        t1 = xr[0] + xr[1]; xr[1] = xr[0] - xr[1]; xr[0] = t1;
        t1 = xi[0] + xi[1]; xi[1] = xi[0] - xi[1]; xi[0] = t1;
        return;
    }
    if (logm == 2) { /* subs.c:202-238 */
        t1 = xr[0] + xr[2]; xr[2] = xr[0] - xr[2]; xr[0] = t1;
        t1 = xi[0] + xi[2]; xi[2] = xi[0] - xi[2]; xi[0] = t1;
        t1 = xr[1] + xr[3]; xr[3] = xr[1] - xr[3]; xr[1] = t1;
        t1 = xi[1] + xi[3]; xi[3] = xi[1] - xi[3]; xi[1] = t1;
        t1 = xr[0] + xr[1]; xr[1] = xr[0] - xr[1]; xr[0] = t1;
        t1 = xi[0] + xi[1]; xi[1] = xi[0] - xi[1]; xi[0] = t1;
        t1 = xr[2] + xi[3];
        t2 = xi[2] + xr[3];
        xi[2] = xi[2] - xr[3];
        xr[3] = xr[2] - xi[3];
        xr[2] = t1;
        xi[3] = t2;
        return;
    }
    m = 1 << logm; m2 = m / 2; m4 = m / 4; m8 = m / 8;
    for (n = 0; n < m2; n++) { /* step 1 */
        t1 = xr[n] + xr[n + m2]; xr[n + m2] = xr[n] - xr[n + m2]; xr[n] = t1;
        t2 = xi[n] + xi[n + m2]; xi[n + m2] = xi[n] - xi[n + m2]; xi[n] = t2;
    }
    for (n = 0; n < m4; n++) { /* step 2 */
        float *a = xr + m2 + n, *b = xr + m2 + m4 + n, *c = xi + m2 + n, *d = xi + m2 + m4 + n;
        t1 = *a + *d;
        t2 = *c + *b;
        *c = *c - *b;
        *b = *a - *d;
        *a = t1;
        *d = t2;
    }
    { /* steps 3&4 */
        int nel = m4 - 2, k = 0;
        const float *t = logm >= 4 ? CT.tw[logm] : NULL;
        for (n = 1; n < m4; n++) {
            float *a = xr + m2 + n, *b = xr + m2 + m4 + n, *c = xi + m2 + n, *d = xi + m2 + m4 + n;
            if (n == m8) {
                t1 = (float)(SQHALF * (*a + *c));
                *c = (float)(SQHALF * (*c - *a));
                *a = t1;
                t2 = (float)(SQHALF * (*d - *b));
                *d = (float)(-SQHALF * (*b + *d));
                *b = t2;
            } else {
                t2 = t[k] * (*a + *c);
                t1 = t[nel + k] * *a + t2;
                *a = t[2 * nel + k] * *c + t2;
                *c = t1;
                t2 = t[3 * nel + k] * (*b + *d);
                t1 = t[4 * nel + k] * *b + t2;
                *b = t[5 * nel + k] * *d + t2;
                *d = t1;
                k++;
            }
        }
    }
    cplx_sr(xr, xi, logm - 1);
    cplx_sr(xr + m2, xi + m2, logm - 2);
    cplx_sr(xr + 3 * m4, xi + 3 * m4, logm - 2);
}

static void real_sr(float *x, int logm) /* rsrec, subs.c:412-523 */
{
    int m, m2, m4, m8, n;
    float t1, t2;
    if (logm <= 0) return;
    if (logm == 1) { t1 = x[0] + x[1]; x[1] = x[0] - x[1]; x[0] = t1; return; }
    m = 1 << logm; m2 = m / 2; m4 = m / 4; m8 = m / 8;
    for (n = 0; n < m2; n++) { t1 = x[n] + x[n + m2]; x[n + m2] = x[n] - x[n + m2]; x[n] = t1; }
    for (n = 0; n < m4; n++) x[m2 + m4 + n] = -x[m2 + m4 + n];
    {
        int nel = m4 - 2, k = 0;
        const float *t = logm >= 4 ? CT.tw[logm] : NULL;
        for (n = 1; n < m4; n++) {
            float *a = x + m2 + n, *c = x + m2 + m4 + n;
            if (n == m8) {
                t1 = (float)(SQHALF * (*a + *c));
                *c = (float)(SQHALF * (*c - *a));
                *a = t1;
            } else {
                t2 = t[k] * (*a + *c);
                t1 = t[nel + k] * *a + t2;
                *a = t[2 * nel + k] * *c + t2;
                *c = t1;
                k++;
            }
        }
    }
    real_sr(x, logm - 1);
    cplx_sr(x + m2, x + 3 * m4, logm - 2);
    { /* step 5, subs.c:501-521 */
        float *p = x + m2 + m4, *q = x + m - 1;
        for (n = 0; n < m8; n++) { t1 = *p; *p++ = -*q; *q-- = -t1; }
        p = x + m2 + 1; q = x + m - 2;
        for (n = 0; n < m8; n++) { t1 = *p; *p = -*q; *q = t1; p += 2; q -= 2; }
    }
    if (logm == 2) x[3] = -x[3];
}

void l3o_fft(float *x, int n, float *energy, float *phi)
{
    int logm = (n == 1024) ? 10 : 8, i, h = n / 2;
    const int *br;
    init_common();
    br = (n == 1024) ? CT.brev10 : CT.brev8;
    real_sr(x, logm);
    for (i = 0; i < n; i++) { int j = br[i]; if (j > i) { float t = x[i]; x[i] = x[j]; x[j] = t; } } /* subs.c:136-177 */
    /* enphinew, subs.c:53-123 */
    energy[0] = x[0] * x[0];
    phi[0] = (float)atan2(0.0, (double)x[0]);
    for (i = 1; i < h; i++) {
        float e = x[i] * x[i] + x[n - i] * x[n - i];
        if (e < 0.0005) { energy[i] = 0.0005f; phi[i] = 0.0f; }
        else { energy[i] = e; phi[i] = (float)atan2(-(double)x[n - i], (double)x[i]); }
    }
    energy[h] = x[h] * x[h];
    phi[h] = (float)atan2(0.0, (double)x[h]);
}

/* ------------------------------------------------------------------------------------------ */
/* psychoacoustic model 2, Layer III branch (l3psy.c:443-740)                                 */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    short savebuf[1344];
    float r[2][6], phi_sav[2][6];
    double nb_1[63], nb_2[63];
    int blocktype_old;
    double ratio[21], ratio_s[12][3];
} psy_chan;

/* 44.1 kHz sparse spreading ranges, l3psy.c:996-1060 (first/last partition per row) */
static const unsigned char SPR_LO[63] = {0,0,0,0,0,0,0,0,0,0,0,1,1,2,3,5,6,7,9,10,11,12,14,15,15,16,16,17,18,19,19,20,
    21,22,22,23,24,25,26,27,28,29,30,31,32,33,34,35,36,37,37,38,39,40,41,42,43,44,45,46,47,48,48};
static const unsigned char SPR_HI[63] = {2,3,4,5,6,7,8,9,10,11,12,14,14,15,15,16,17,19,20,21,22,23,24,25,27,28,28,29,30,31,32,34,
    35,36,36,37,38,39,41,42,43,44,45,46,47,48,49,50,51,52,53,54,55,56,57,58,59,60,61,62,62,62,62};

static double unpred(double r_new, double phi_new, double r_prime, double phi_prime) /* l3psy.c:503-511 */
{
    double t1 = r_new * cos(phi_new) - r_prime * cos(phi_prime);
    double t2 = r_new * sin(phi_new) - r_prime * sin(phi_prime);
    double t3 = r_new + fabs(r_prime);
    if (t3 != 0.0) return sqrt(t1 * t1 + t2 * t2) / t3;
    return 0.0;
}

/* one L3psycho_anal call; cur = index of the "new" history slot (toggled by the caller on ch 0) */
static void psy_granule(const psy_tables *T, psy_chan *S, int cur, const short *pcm576,
                        double ratio_d[21], double ratio_ds[12][3], double *pe_out, int *block_type_out)
{
    float wsamp[1024], energy[513], phi[513], energy_s[3][129], phi_s[3][129];
    float cb[63], ecb[63], nb[63];
    double cw[513], eb[63], ctb[63], thr[63], snr_l[63], en[21], thm[21];
    int old = 1 - cur, oldest = cur;
    int b, j, k, sb, sblock, blocktype;
    double pe;

    memcpy(ratio_d, S->ratio, sizeof(S->ratio));                 /* l3psy.c:452-456 */
    memcpy(ratio_ds, S->ratio_s, sizeof(S->ratio_s));
    memmove(S->savebuf, S->savebuf + 576, 768 * sizeof(short));  /* l3psy.c:477-481 */
    memcpy(S->savebuf + 768, pcm576, 576 * sizeof(short));

    for (j = 0; j < 1024; j++) wsamp[j] = CT.hann_l[j] * S->savebuf[j];
    l3o_fft(wsamp, 1024, energy, phi);
    for (j = 0; j < 6; j++) {                                    /* l3psy.c:496-512 */
        double r_prime = 2.0 * S->r[old][j] - S->r[oldest][j];
        double phi_prime = 2.0 * S->phi_sav[old][j] - S->phi_sav[oldest][j];
        S->r[cur][j] = (float)sqrt((double)energy[j]);
        S->phi_sav[cur][j] = phi[j];
        cw[j] = unpred(S->r[cur][j], (double)phi[j], r_prime, phi_prime);
    }
    for (sblock = 0; sblock < 3; sblock++) {                     /* l3psy.c:518-527 */
        for (j = 0, k = 128 * (2 + sblock); j < 256; j++, k++) wsamp[j] = CT.hann_s[j] * S->savebuf[k];
        l3o_fft(wsamp, 256, energy_s[sblock], phi_s[sblock]);
    }
    for (j = 6; j < 206; j += 4) {                               /* l3psy.c:531-549 */
        double r_prime, phi_prime, r2, phi2;
        k = (j + 2) >> 2;
        r_prime = 2.0 * sqrt((double)energy_s[0][k]) - sqrt((double)energy_s[2][k]);
        phi_prime = 2.0 * phi_s[0][k] - phi_s[2][k];
        r2 = sqrt((double)energy_s[1][k]);
        phi2 = phi_s[1][k];
        cw[j] = unpred(r2, phi2, r_prime, phi_prime);
        cw[j + 1] = cw[j + 2] = cw[j + 3] = cw[j];
    }
    for (j = 206; j < 513; j++) cw[j] = 0.4;                     /* l3psy.c:555-556 */

    for (b = 0; b < 63; b++) { eb[b] = 0.0; cb[b] = 0.0f; }      /* l3psy.c:565-578 */
    for (j = 0; j < 513; j++) {
        int tp = T->part_l[j];
        eb[tp] += energy[j];
        cb[tp] = (float)(cb[tp] + cw[j] * energy[j]);
    }
    for (b = 0; b < 63; b++) { ecb[b] = 0.0f; ctb[b] = 0.0; }    /* l3psy.c:586-605 */
    if (T->sr_idx == 1) {
        for (b = 0; b < 63; b++)
            for (k = SPR_LO[b]; k <= SPR_HI[b]; k++) ecb[b] = (float)(ecb[b] + T->s3_l[b][k] * eb[k]);
        for (b = 0; b < 63; b++)
            for (k = SPR_LO[b]; k <= SPR_HI[b]; k++) ctb[b] += T->s3_l[b][k] * cb[k];
    } else {
        for (b = 0; b < 63; b++)
            for (k = 0; k < 63; k++)
                if (T->s3_l[b][k] != 1.0) {
                    ecb[b] = (float)(ecb[b] + T->s3_l[b][k] * eb[k]);
                    ctb[b] += T->s3_l[b][k] * cb[k];
                }
    }
    for (b = 0; b < 63; b++) {                                   /* l3psy.c:610-624 */
        double cbb, tbb, v;
        if (ecb[b] != 0.0) {
            cbb = ctb[b] / ecb[b];
            if (cbb < 0.01) cbb = 0.01;
            cbb = log(cbb);
        } else cbb = 0.0;
        tbb = -0.299 - 0.43 * cbb;
        tbb = (0.0 > tbb) ? 0.0 : tbb;
        tbb = (1.0 < tbb) ? 1.0 : tbb;
        v = 29.0 * tbb + 6.0 * (1.0 - tbb);
        snr_l[b] = (T->minval[b] > v) ? T->minval[b] : v;
    }
    for (b = 0; b < 63; b++) nb[b] = (float)(ecb[b] * T->norm_l[b] * exp(-snr_l[b] * LN_TO_LOG10)); /* :626-627 */
    for (b = 0; b < 63; b++) {                                   /* l3psy.c:629-636 */
        double a = 2.0 * S->nb_1[b], c = 16.0 * S->nb_2[b];
        double mn = (a < c) ? a : c;
        double t = (nb[b] < mn) ? (double)nb[b] : mn;
        thr[b] = (T->qthr_l[b] > t) ? T->qthr_l[b] : t;
        S->nb_2[b] = S->nb_1[b];
        S->nb_1[b] = nb[b];
    }
    pe = 0.0;                                                    /* l3psy.c:639-645 */
    for (b = 0; b < 63; b++) {
        double l = log((thr[b] + 1.0) / (eb[b] + 1.0));
        double tp = (0.0 < l) ? 0.0 : l;
        pe -= T->numlines_pe[b] * tp;
    }
    *pe_out = pe;

    blocktype = 0;
    if (pe < 1800) {                                             /* l3psy.c:651-685 */
        if (S->blocktype_old == 2) blocktype = 3;                /* SHORT -> STOP */
        else blocktype = 0;
        for (sb = 0; sb < 21; sb++) {
            en[sb] = T->w1_l[sb] * eb[T->bu_l[sb]] + T->w2_l[sb] * eb[T->bo_l[sb]];
            thm[sb] = T->w1_l[sb] * thr[T->bu_l[sb]] + T->w2_l[sb] * thr[T->bo_l[sb]];
            for (b = T->bu_l[sb] + 1; b < T->bo_l[sb]; b++) { en[sb] += eb[b]; thm[sb] += thr[b]; }
            S->ratio[sb] = (en[sb] != 0.0) ? thm[sb] / en[sb] : 0.0;
        }
    } else {                                                     /* l3psy.c:686-730 */
        blocktype = 2;
        if (S->blocktype_old == 0) S->blocktype_old = 1;
        if (S->blocktype_old == 3) S->blocktype_old = 2;
        for (sblock = 0; sblock < 3; sblock++) {
            for (b = 0; b < 42; b++) { eb[b] = 0.0; ecb[b] = 0.0f; }
            for (j = 0; j < 129; j++) eb[T->part_s[j]] += energy_s[sblock][j];
            for (b = 0; b < 42; b++)
                for (k = 0; k < 42; k++) ecb[b] = (float)(ecb[b] + T->s3_l[b][k] * eb[k]);
            for (b = 0; b < 42; b++) {
                nb[b] = (float)(ecb[b] * T->norm_l[b] * exp((double)T->snr_s[b] * LN_TO_LOG10));
                thr[b] = (T->qthr_s[b] > nb[b]) ? T->qthr_s[b] : (double)nb[b];
            }
            for (sb = 0; sb < 12; sb++) {
                double e = T->w1_s[sb] * eb[T->bu_s[sb]] + T->w2_s[sb] * eb[T->bo_s[sb]];
                double t = T->w1_s[sb] * thr[T->bu_s[sb]] + T->w2_s[sb] * thr[T->bo_s[sb]];
                for (b = T->bu_s[sb] + 1; b < T->bo_s[sb]; b++) { e += eb[b]; t += thr[b]; }
                S->ratio_s[sb][sblock] = (e != 0.0) ? t / e : 0.0;
            }
        }
    }
    *block_type_out = S->blocktype_old;                          /* l3psy.c:732-739 */
    S->blocktype_old = blocktype;
}

/* ------------------------------------------------------------------------------------------ */
/* rate loop (loop.c) + bit reservoir (reservoir.c)                                           */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    l3o_gr_info g;
    int sfb_lmax, sfb_smax;
    double q;  /* quantizerStepSize */
} work_gi;

typedef struct {
    int resv_size, resv_max;
    int en_tot[2][2], en[2][2][21], xm[2][2][21], xrmax[2][2]; /* calc_scfsi statics, loop.c:618-621 */
    int addr[2][2][3]; /* address1..3 live in main()'s static l3_side and are NOT reset per frame:
                          subdivide() leaves them stale when big_values == 0 (loop.c:1642-1647) */
} loop_state;

static const int PRETAB[21] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 3, 2}; /* Table B.6 */
static const int SCFSI_BAND[5] = {0, 6, 11, 16, 21};
static const int SLEN1[16] = {0, 0, 0, 0, 3, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4};
static const int SLEN2[16] = {0, 1, 2, 3, 0, 1, 2, 3, 1, 2, 3, 1, 2, 3, 2, 3};
static const unsigned char SUBDV[23][2] = {{0,0},{0,0},{0,0},{0,0},{0,0},{0,1},{1,1},{1,1},{1,2},{2,2},{2,3},{2,3},
    {3,4},{3,4},{3,4},{4,5},{4,5},{4,6},{5,6},{5,6},{5,7},{6,7},{6,7}}; /* loop.c:1596-1625 */

static int nint_ref(double in) { return (in < 0) ? (int)(in - 0.5) : (int)(in + 0.5); } /* loop.c:2021-2030 */

static int pow_nint_ref(double x) /* pow_nint.h:16-50 */
{
    int step = 1, pos = 1, p = 0;
    while (pos < 2048) {
        if (x < CT.pow_nint_tab[pos]) break;
        p = pos; pos += step; step <<= 1;
    }
    step >>= 1; pos -= step; step >>= 1;
    if (step) {
        while (step) {
            if (x < CT.pow_nint_tab[pos]) pos -= step;
            else { p = pos; pos += step; }
            step >>= 1;
        }
        if (x >= CT.pow_nint_tab[pos]) p = pos;
    }
    return p;
}

static void quantize_ref(const double *xr, int *ix, double q) /* loop.c:1360-1428 (subblock_gain==0, no mixed blocks) */
{
    double step = (q == 0.0) ? 1.0 : pow(2.0, q * 0.25);
    double ostep = 1.0 / step;
    int i;
    for (i = 0; i < 576; i++) ix[i] = pow_nint_ref(fabs(xr[i]) * ostep);
}

static int hlen_of(int table, int x, int y) { return MP3T_HLEN[MP3T_HUFF[table].off + x * MP3T_HUFF[table].ylen + y]; }

static int count_bit_ref(const int *ix, unsigned start, unsigned end, unsigned table) /* loop.c:172-225 */
{
    unsigned i;
    int sum = 0, linbits;
    if (table == 0) return 0;
    linbits = MP3T_HUFF[table].linbits;
    for (i = start; i < end; i += 2) {
        int x = ix[i], y = ix[i + 1];
        if (table > 15) {
            if (x > 14) { x = 15; sum += linbits; }
            if (y > 14) { y = 15; sum += linbits; }
        }
        sum += hlen_of((int)table, x, y);
        if (x != 0) sum++;
        if (y != 0) sum++;
    }
    return sum;
}

static int ix_max_ref(const int *ix, unsigned b, unsigned e)
{
    unsigned i; int mx = 0;
    for (i = b; i < e; i++) if (ix[i] > mx) mx = ix[i];
    return mx;
}

static int is_short(const l3o_gr_info *g) { return g->window_switching_flag && g->block_type == 2; }

static void calc_runlen_ref(const int *ix, l3o_gr_info *g) /* loop.c:1488-1519 */
{
    int i;
    if (is_short(g)) { g->count1 = 0; g->big_values = 288; return; }
    for (i = 576; i > 1; i -= 2) if (!(ix[i - 1] == 0 && ix[i - 2] == 0)) break;
    g->count1 = 0;
    for (; i > 3; i -= 4) {
        if (ix[i - 1] <= 1 && ix[i - 2] <= 1 && ix[i - 3] <= 1 && ix[i - 4] <= 1) g->count1++;
        else break;
    }
    g->big_values = i / 2;
}

static int count1_bitcount_ref(const int *ix, l3o_gr_info *g) /* loop.c:1531-1590 */
{
    int i, k, sum0 = 0, sum1 = 0;
    for (i = g->big_values * 2, k = 0; k < g->count1; i += 4, k++) {
        int v = ix[i], w = ix[i + 1], x = ix[i + 2], y = ix[i + 3];
        int p = v + (w << 1) + (x << 2) + (y << 3);
        int sign = (v != 0) + (w != 0) + (x != 0) + (y != 0);
        sum0 += sign + MP3T_HLEN[MP3T_HUFF[32].off + p];
        sum1 += sign + MP3T_HLEN[MP3T_HUFF[33].off + p];
    }
    if (sum0 < sum1) { g->count1table_select = 0; return sum0; }
    g->count1table_select = 1;
    return sum1;
}

static void subdivide_ref(l3o_gr_info *g, const short *sfb_l) /* loop.c:1638-1704 */
{
    if (g->big_values == 0) { g->region0_count = 0; g->region1_count = 0; return; }
    {
        int bvr = 2 * g->big_values;
        if (g->window_switching_flag == 0) {
            int n = 0, thiscount, index;
            while (sfb_l[n] < bvr) n++;
            thiscount = SUBDV[n][0]; index = thiscount + 1;
            while (thiscount && sfb_l[index] > bvr) { thiscount--; index--; }
            g->region0_count = thiscount;
            thiscount = SUBDV[n][1];
            index = g->region0_count + thiscount + 2;
            while (thiscount && sfb_l[index] > bvr) { thiscount--; index--; }
            g->region1_count = thiscount;
            g->address1 = sfb_l[g->region0_count + 1];
            g->address2 = sfb_l[g->region0_count + g->region1_count + 2];
            g->address3 = bvr;
        } else if (g->block_type == 2 && g->mixed_block_flag == 0) {
            g->region0_count = 8; g->region1_count = 36;
            g->address1 = 36; g->address2 = bvr; g->address3 = 0;
        } else {
            g->region0_count = 7; g->region1_count = 13;
            g->address1 = sfb_l[8]; g->address2 = bvr; g->address3 = 0;
        }
    }
}

static int choose_table_ref(int max) /* loop.c:1908-1943 */
{
    int i;
    if (max == 0) return 0;
    if (max < 15) { for (i = 0; i < 15; i++) if (MP3T_HUFF[i].xlen > max) return i; }
    else { max -= 15; for (i = 15; i < 32; i++) if (MP3T_HUFF[i].linmax >= max) return i; }
    return 0;
}

static int new_choose_table_ref(const int *ix, unsigned begin, unsigned end) /* loop.c:1793-1900 */
{
    int i, max = ix_max_ref(ix, begin, end), c0 = 0, c1 = 0, s0, s1;
    if (max == 0) return 0;
    if (max < 15) {
        for (i = 0; i < 14; i++) if (MP3T_HUFF[i].xlen > max) { c0 = i; break; }
        s0 = count_bit_ref(ix, begin, end, c0);
        switch (c0) {
        case 2: s1 = count_bit_ref(ix, begin, end, 3); if (s1 <= s0) c0 = 3; break;
        case 5: s1 = count_bit_ref(ix, begin, end, 6); if (s1 <= s0) c0 = 6; break;
        case 7:
            s1 = count_bit_ref(ix, begin, end, 8); if (s1 <= s0) { c0 = 8; s0 = s1; }
            s1 = count_bit_ref(ix, begin, end, 9); if (s1 <= s0) c0 = 9;
            break;
        case 10:
            s1 = count_bit_ref(ix, begin, end, 11); if (s1 <= s0) { c0 = 11; s0 = s1; }
            s1 = count_bit_ref(ix, begin, end, 12); if (s1 <= s0) c0 = 12;
            break;
        case 13: s1 = count_bit_ref(ix, begin, end, 15); if (s1 <= s0) c0 = 15; break;
        default: break;
        }
    } else {
        max -= 15;
        for (i = 15; i < 24; i++) if (MP3T_HUFF[i].linmax >= max) { c0 = i; break; }
        for (i = 24; i < 32; i++) if (MP3T_HUFF[i].linmax >= max) { c1 = i; break; }
        s0 = count_bit_ref(ix, begin, end, c0);
        s1 = count_bit_ref(ix, begin, end, c1);
        if (s1 < s0) c0 = c1;
    }
    return c0;
}

static void bigv_tab_select_ref(const int *ix, l3o_gr_info *g, const short *sfb_s) /* loop.c:1717-1775 */
{
    g->table_select[0] = g->table_select[1] = g->table_select[2] = 0;
    if (is_short(g)) {
        int sfb, w, line, max1 = 0, max2 = 0;
        for (sfb = 0; sfb < 13; sfb++) {
            int start = sfb_s[sfb], end = sfb_s[sfb + 1];
            int *pm = (start < 12) ? &max1 : &max2;
            for (w = 0; w < 3; w++)
                for (line = start; line < end; line += 2) {
                    int x = ix[line * 3 + w], y = ix[(line + 1) * 3 + w];
                    if (x > *pm) *pm = x;
                    if (y > *pm) *pm = y;
                }
        }
        g->table_select[0] = choose_table_ref(max1);
        g->table_select[1] = choose_table_ref(max2);
    } else {
        if (g->address1 > 0) g->table_select[0] = new_choose_table_ref(ix, 0, g->address1);
        if (g->address2 > g->address1) g->table_select[1] = new_choose_table_ref(ix, g->address1, g->address2);
        if (g->big_values * 2 > g->address2) g->table_select[2] = new_choose_table_ref(ix, g->address2, g->big_values * 2);
    }
}

static int bigv_bitcount_ref(const int *ix, const l3o_gr_info *g, const short *sfb_s) /* loop.c:1954-2016 */
{
    int bits = 0;
    if (is_short(g)) {
        int sfb, w, line;
        for (sfb = 0; sfb < 13; sfb++) {
            int start = sfb_s[sfb], end = sfb_s[sfb + 1];
            int t = (start < 12) ? g->table_select[0] : g->table_select[1];
            for (w = 0; w < 3; w++)
                for (line = start; line < end; line += 2) {
                    int pair[2];
                    pair[0] = ix[line * 3 + w]; pair[1] = ix[(line + 1) * 3 + w];
                    bits += count_bit_ref(pair, 0, 2, t); /* HuffmanCode count mode, huffcode.h:15-139 */
                }
        }
    } else {
        if (g->table_select[0]) bits += count_bit_ref(ix, 0, g->address1, g->table_select[0]);
        if (g->table_select[1]) bits += count_bit_ref(ix, g->address1, g->address2, g->table_select[1]);
        if (g->table_select[2]) bits += count_bit_ref(ix, g->address2, g->address3, g->table_select[2]);
    }
    return bits;
}

static int count_bits_ref(const int *ix, l3o_gr_info *g, int sr) /* loop.c:2099-2113 == inner_loop body :590-594 */
{
    int bits;
    calc_runlen_ref(ix, g);
    bits = count1_bitcount_ref(ix, g);
    subdivide_ref(g, MP3T_SFB_LONG[sr]);
    bigv_tab_select_ref(ix, g, MP3T_SFB_SHORT[sr]);
    bits += bigv_bitcount_ref(ix, g, MP3T_SFB_SHORT[sr]);
    return bits;
}

static void set_block(l3o_gr_info *g, int block_type)
{
    g->block_type = block_type;
    g->window_switching_flag = (block_type != 0);
    g->mixed_block_flag = 0;
}

int l3o_count_bits(const int *ix, int block_type, int sr_idx, l3o_gr_info *gi)
{
    init_common();
    set_block(gi, block_type);
    return count_bits_ref(ix, gi, sr_idx);
}

int l3o_quantize_count(const double *xr_abs, int q, int block_type, int sr_idx, int *ix, l3o_gr_info *gi)
{
    init_common();
    set_block(gi, block_type);
    quantize_ref(xr_abs, ix, (double)q);
    return count_bits_ref(ix, gi, sr_idx);
}

static int part2_length_ref(const l3o_gr_info *g, int gr, const int scfsi[4]) /* loop.c:731-784 (MPEG-1) */
{
    int s1 = SLEN1[g->scalefac_compress], s2 = SLEN2[g->scalefac_compress], bits = 0;
    if (g->window_switching_flag == 1 && g->block_type == 2) bits += 18 * s1 + 18 * s2;
    else {
        if (gr == 0 || scfsi[0] == 0) bits += 6 * s1;
        if (gr == 0 || scfsi[1] == 0) bits += 5 * s1;
        if (gr == 0 || scfsi[2] == 0) bits += 5 * s2;
        if (gr == 0 || scfsi[3] == 0) bits += 5 * s2;
    }
    return bits;
}

static int scale_bitcount_ref(const int sf_l[22], int sf_s[13][3], l3o_gr_info *g) /* loop.c:792-857 */
{
    static const int pow2[5] = {1, 2, 4, 8, 16};
    int i, k, sfb, m1 = 0, m2 = 0, ep = 2;
    if (is_short(g)) {
        for (i = 0; i < 3; i++) {
            for (sfb = 0; sfb < 6; sfb++) if (sf_s[sfb][i] > m1) m1 = sf_s[sfb][i];
            for (sfb = 6; sfb < 12; sfb++) if (sf_s[sfb][i] > m2) m2 = sf_s[sfb][i];
        }
    } else {
        for (sfb = 0; sfb < 11; sfb++) if (sf_l[sfb] > m1) m1 = sf_l[sfb];
        for (sfb = 11; sfb < 21; sfb++) if (sf_l[sfb] > m2) m2 = sf_l[sfb];
    }
    for (k = 0; k < 16; k++)
        if (m1 < pow2[SLEN1[k]] && m2 < pow2[SLEN2[k]]) { ep = 0; break; }
    if (ep == 0) g->scalefac_compress = k;
    return ep;
}

static void calc_noise_ref(const double *xr, const int *ix, const work_gi *w, int sr, double xfsf[4][21]) /* loop.c:1007-1069 */
{
    const short *sl = MP3T_SFB_LONG[sr], *ss = MP3T_SFB_SHORT[sr];
    double step = pow(2.0, w->q * 0.25);
    int sfb, l, i;
    for (sfb = 0; sfb < w->sfb_lmax; sfb++) {
        double sum = 0.0, bw = sl[sfb + 1] - sl[sfb];
        for (l = sl[sfb]; l < sl[sfb + 1]; l++) {
            double t = fabs(xr[l]) - CT.pow43[ix[l]] * step;
            sum += t * t;
        }
        xfsf[0][sfb] = sum / bw;
    }
    for (i = 0; i < 3; i++)
        for (sfb = w->sfb_smax; sfb < 12; sfb++) {
            double sum = 0.0, bw = ss[sfb + 1] - ss[sfb];
            for (l = ss[sfb]; l < ss[sfb + 1]; l++) {
                double t = fabs(xr[l * 3 + i]) - CT.pow43[ix[l * 3 + i]] * step;
                sum += t * t;
            }
            xfsf[i + 1][sfb] = sum / bw;
        }
}

typedef struct { double l[21]; double s[12][3]; } xmin_t;

static void calc_xmin_ref(const double *xr, const double ratio_l[21], double ratio_s[12][3], const work_gi *w, int sr, xmin_t *xm) /* loop.c:1085-1119 */
{
    const short *sl = MP3T_SFB_LONG[sr], *ss = MP3T_SFB_SHORT[sr];
    int sfb, l, b;
    for (sfb = w->sfb_smax; sfb < 12; sfb++) {
        double bw = ss[sfb + 1] - ss[sfb];
        for (b = 0; b < 3; b++) {
            double en = 0.0;
            for (l = ss[sfb]; l < ss[sfb + 1]; l++) en += xr[l * 3 + b] * xr[l * 3 + b];
            xm->s[sfb][b] = ratio_s[sfb][b] * en / bw;
        }
    }
    for (sfb = 0; sfb < w->sfb_lmax; sfb++) {
        double bw = sl[sfb + 1] - sl[sfb], en = 0.0;
        for (l = sl[sfb]; l < sl[sfb + 1]; l++) en += xr[l] * xr[l];
        xm->l[sfb] = ratio_l[sfb] * en / bw;
    }
}

/* loop.c:615-722, int-typed statics and swapped indices reproduced on purpose */
static void calc_scfsi_ref(loop_state *S, const double *xr, const work_gi *w, const xmin_t *xm, int sr, int ch, int gr, int scfsi_ch[4])
{
    const short *sl = MP3T_SFB_LONG[sr];
    double temp, log2v = log(2.0), mx = 0.0;
    int sfb, i, condition = 0;
    for (i = 0; i < 576; i++) { double a = fabs(xr[i]); if (a > mx) mx = a; }
    S->xrmax[gr][ch] = (int)mx;
    for (temp = 0.0, i = 0; i < 576; i++) temp += xr[i] * xr[i];
    S->en_tot[gr][ch] = (temp == 0.0) ? 0 : (int)(log(temp) / log2v);
    if (w->g.window_switching_flag == 0 || w->g.block_type != 2)
        for (sfb = 0; sfb < 21; sfb++) {
            for (temp = 0.0, i = sl[sfb]; i < sl[sfb + 1]; i++) temp += xr[i] * xr[i];
            S->en[gr][ch][sfb] = (temp == 0.0) ? 0 : (int)(log(temp) / log2v);
            S->xm[gr][ch][sfb] = (xm->l[sfb] == 0.0) ? 0 : (int)(log(xm->l[sfb]) / log2v);
        }
    if (gr == 1) {
        int gr2, tp, band;
        for (gr2 = 0; gr2 < 2; gr2++) {
            if (S->xrmax[ch][gr2] != 0) condition++;
            if (w->g.window_switching_flag == 0 || w->g.block_type != 2) condition++;
        }
        condition++; /* abs(en_tot[0] - en_tot[1]) is a pointer difference == 2 < 10, loop.c:683 */
        for (tp = 0, sfb = 0; sfb < 21; sfb++) tp += abs(S->en[ch][0][sfb] - S->en[ch][1][sfb]);
        if (tp < 100) condition++;
        if (condition == 6) {
            for (band = 0; band < 4; band++) {
                int sum0 = 0, sum1 = 0;
                for (sfb = SCFSI_BAND[band]; sfb < SCFSI_BAND[band + 1]; sfb++) {
                    sum0 += abs(S->en[ch][0][sfb] - S->en[ch][1][sfb]);
                    sum1 += abs(S->xm[ch][0][sfb] - S->xm[ch][1][sfb]);
                }
                scfsi_ch[band] = (sum0 < 10 && sum1 < 10) ? 1 : 0;
            }
        } else for (band = 0; band < 4; band++) scfsi_ch[band] = 0;
    }
}

static int quantanf_init_ref(const double *xr) /* loop.c:369-402 */
{
    int i, tp = 0;
    double sum1 = 0.0, sum2 = 0.0;
    for (i = 0; i < 576; i++)
        if (xr[i] != 0) { double t = xr[i] * xr[i]; sum1 += log(t); sum2 += t; }
    if (sum2 != 0.0) {
        double sfm = exp(sum1 / 576.0) / (sum2 / 576.0);
        tp = nint_ref(8.0 * log(sfm));
        if (tp < -100.0) tp = -100;
    }
    return (int)(tp - 70.0);
}

static int resv_max_bits(const loop_state *S, double pe, int mean_bits, int n_ch) /* reservoir.c:101-134 */
{
    int more_bits, max_bits, add_bits = 0, over_bits;
    mean_bits /= n_ch;
    max_bits = mean_bits;
    if (max_bits > 4095) max_bits = 4095;
    if (S->resv_max == 0) return max_bits;
    more_bits = (int)(pe * 3.1 - mean_bits);
    if (more_bits > 100) {
        int frac = (S->resv_size * 6) / 10;
        add_bits = (frac < more_bits) ? frac : more_bits;
    }
    over_bits = S->resv_size - ((S->resv_max * 8) / 10) - add_bits;
    if (over_bits > 0) add_bits += over_bits;
    max_bits += add_bits;
    if (max_bits > 4095) max_bits = 4095;
    return max_bits;
}

/* outer_loop + inner_loop + bin_search_StepSize, loop.c:415-606, 2119-2140. xr is mutated. */
static int outer_loop_ref(double *xr, int max_bits, xmin_t *xm, int *ix, work_gi *w, int sf_l[22], int sf_s[13][3],
                          int gr, int sr, const int scfsi_ch[4], const l3o_gr_info *gr0, const int gr0_sf_l[22])
{
    int save_l[21], save_s[13][3], save_preflag = 0, save_compress = 0;
    int status, over, iteration = 0, bits = 0, sfb, i, l;
    double xfsf[4][21];
    const short *sl = MP3T_SFB_LONG[sr], *ss = MP3T_SFB_SHORT[sr];
    memset(xfsf, 0, sizeof(xfsf));
    do {
        int huff_bits;
        iteration++;
        w->g.part2_length = part2_length_ref(&w->g, gr, scfsi_ch);
        huff_bits = max_bits - w->g.part2_length;
        if (iteration == 1) { /* bin_search_StepSize(max_bits, ...) */
            double top = w->q, bot = 200, next = w->q, last;
            int bit;
            do {
                last = next;
                next = (double)(long)((top + bot) / 2.0);
                w->q = next;
                quantize_ref(xr, ix, w->q);
                bit = count_bits_ref(ix, &w->g, sr);
                if (bit > max_bits) top = next; else bot = next;
            } while (bit != max_bits && fabs(last - next) > 1.0);
        }
        /* inner_loop(huff_bits) */
        w->q -= 1.0;
        do {
            w->q += 1.0;
            quantize_ref(xr, ix, w->q);
            bits = count_bits_ref(ix, &w->g, sr);
        } while (bits > huff_bits && w->q < 1024); /* guard: the reference assert()s huff_bits >= 0, loop.c:579 */

        calc_noise_ref(xr, ix, w, sr, xfsf);
        for (sfb = 0; sfb < 21; sfb++) save_l[sfb] = sf_l[sfb];
        memcpy(save_s, sf_s, sizeof(save_s));
        save_preflag = w->g.preflag;
        save_compress = w->g.scalefac_compress;

        { /* preemphasis, loop.c:1161-1216 */
            int done = 0, band;
            if (gr == 1)
                for (band = 0; band < 4; band++)
                    if (scfsi_ch[band]) { w->g.preflag = gr0->preflag; done = 1; break; }
            if (!done && w->g.block_type != 2 && w->g.preflag == 0) {
                int ov = 0;
                for (sfb = 17; sfb < 21; sfb++) if (xfsf[0][sfb] > xm->l[sfb]) ov++;
                if (ov == 4) {
                    double ifq = sqrt(2.);
                    w->g.preflag = 1;
                    for (sfb = 0; sfb < w->sfb_lmax; sfb++) {
                        xm->l[sfb] *= pow(ifq, 2.0 * (double)PRETAB[sfb]);
                        for (i = sl[sfb]; i < sl[sfb + 1]; i++) xr[i] *= pow(ifq, (double)PRETAB[sfb]);
                    }
                }
            }
        }
        { /* amp_scalefac_bands, loop.c:1225-1349 */
            double ifq = sqrt(2.0), ifq2;
            int copySF = 0, preventSF = 0, band = 0, sb;
            over = 0;
            if (gr == 1)
                for (sb = 0; sb < 4; sb++)
                    if (scfsi_ch[sb]) {
                        ifq = (gr0->scalefac_scale == 0) ? sqrt(2.0) : pow(2.0, 0.5 * (1.0 + (double)gr0->scalefac_scale));
                        if (iteration == 1) copySF = 1; else preventSF = 1;
                        break;
                    }
            ifq2 = ifq * ifq;
            for (sfb = 0; sfb < w->sfb_lmax; sfb++) {
                if (copySF || preventSF) {
                    if (sfb == SCFSI_BAND[band + 1]) band++;
                    if (scfsi_ch[band]) {
                        if (copySF) sf_l[sfb] = gr0_sf_l[sfb];
                        continue;
                    }
                }
                if (xfsf[0][sfb] > xm->l[sfb]) {
                    over++;
                    xm->l[sfb] *= ifq2;
                    sf_l[sfb]++;
                    for (l = sl[sfb]; l < sl[sfb + 1]; l++) xr[l] *= ifq;
                }
            }
            for (i = 0; i < 3; i++)
                for (sfb = w->sfb_smax; sfb < 12; sfb++)
                    if (xfsf[i + 1][sfb] > xm->s[sfb][i]) {
                        over++;
                        xm->s[sfb][i] *= ifq2;
                        sf_s[sfb][i]++;
                        for (l = ss[sfb]; l < ss[sfb + 1]; l++) xr[l * 3 + i] *= ifq;
                    }
        }
        { /* loop_break, loop.c:1131-1150 */
            status = 1;
            for (sfb = 0; sfb < w->sfb_lmax; sfb++) if (sf_l[sfb] == 0) status = 0;
            for (sfb = w->sfb_smax; sfb < 12; sfb++) for (i = 0; i < 3; i++) if (sf_s[sfb][i] == 0) status = 0;
        }
        if (status == 0) status = scale_bitcount_ref(sf_l, sf_s, &w->g);
    } while (status == 0 && over > 0);

    w->g.preflag = save_preflag;
    w->g.scalefac_compress = save_compress;
    for (sfb = 0; sfb < 21; sfb++) sf_l[sfb] = save_l[sfb];
    for (i = 0; i < 3; i++) for (sfb = 0; sfb < 12; sfb++) sf_s[sfb][i] = save_s[sfb][i];
    w->g.part2_length = part2_length_ref(&w->g, gr, scfsi_ch);
    w->g.part2_3_length = w->g.part2_length + bits;
    return w->g.part2_3_length;
}

/* ------------------------------------------------------------------------------------------ */
/* stream encoder                                                                             */
/* ------------------------------------------------------------------------------------------ */
struct l3o_enc {
    int sfreq, sr_idx, n_ch, bitrate;
    int bits_per_frame, mean_bits;
    double hist[2][512];         /* filterbank history */
    double sb_prev[2][576];      /* sign-fixed previous granule, mdct.c:99-102 */
    psy_chan psy[2];
    int psy_new;                 /* l3psy.c:83 'new' */
    loop_state loop;
};

l3o_enc *l3o_create(int sfreq, int n_ch, int bitrate_kbps)
{
    l3o_enc *e;
    int sr = (sfreq == 32000) ? 0 : (sfreq == 44100) ? 1 : (sfreq == 48000) ? 2 : -1;
    if (sr < 0 || n_ch < 1 || n_ch > 2) return NULL;
    init_common();
    init_psy(sr);
    e = (l3o_enc *)calloc(1, sizeof(*e));
    e->sfreq = sfreq; e->sr_idx = sr; e->n_ch = n_ch; e->bitrate = bitrate_kbps;
    { /* musicin.c:562-572, 729-746 */
        double avg = (1152.0 / (sfreq / 1000.0)) * ((double)bitrate_kbps / 8.0);
        int whole = (int)avg;
        e->bits_per_frame = 8 * whole;
        e->mean_bits = (e->bits_per_frame - (32 + (n_ch == 1 ? 136 : 256))) / 2;
    }
    e->psy_new = 0;
    return e;
}

void l3o_destroy(l3o_enc *e) { free(e); }

int l3o_encode_frame(l3o_enc *e, const short *pcm, l3o_frame *out)
{
    int gr, ch, j, i, n_ch = e->n_ch, sr = e->sr_idx;
    loop_state *L = &e->loop;
    work_gi wg[2][2];
    double cur[576];
    memset(out, 0, sizeof(*out));
    /* psychoacoustics first, musicin.c:751-758 */
    for (gr = 0; gr < 2; gr++)
        for (ch = 0; ch < n_ch; ch++) {
            if (ch == 0) e->psy_new = 1 - e->psy_new;            /* l3psy.c:458-470 */
            psy_granule(&PT[sr], &e->psy[ch], e->psy_new, pcm + ch * 1152 + gr * 576,
                        out->ratio_l[gr][ch], out->ratio_s[gr][ch], &out->pe[gr][ch], &out->block_type[gr][ch]);
        }
    /* polyphase, musicin.c:763-769 */
    for (gr = 0; gr < 2; gr++)
        for (ch = 0; ch < n_ch; ch++)
            for (j = 0; j < 18; j++)
                polyphase_slot(e->hist[ch], pcm + ch * 1152 + gr * 576 + 32 * j, out->sb[gr][ch][j]);
    /* mdct, musicin.c:774 */
    for (gr = 0; gr < 2; gr++)
        for (ch = 0; ch < n_ch; ch++) {
            sign_fix(&out->sb[gr][ch][0][0], cur);
            mdct_granule_fixed(e->sb_prev[ch], cur, out->block_type[gr][ch], out->xr[gr][ch]);
            memcpy(e->sb_prev[ch], cur, sizeof(cur));
        }
    /* iteration_loop, loop.c:232-362 */
    L->resv_max = (e->bits_per_frame > 7680) ? 0 : 7680 - e->bits_per_frame;  /* reservoir.c:45-91 */
    if (L->resv_max > 4088) L->resv_max = 4088;
    for (gr = 0; gr < 2; gr++)
        for (ch = 0; ch < n_ch; ch++) {
            work_gi *w = &wg[gr][ch];
            double xr[576], mx = 0.0;
            xmin_t xm;
            int max_bits;
            memset(w, 0, sizeof(*w));
            memset(&xm, 0, sizeof(xm));
            set_block(&w->g, out->block_type[gr][ch]);
            w->g.address1 = L->addr[gr][ch][0]; w->g.address2 = L->addr[gr][ch][1]; w->g.address3 = L->addr[gr][ch][2];
            if (is_short(&w->g)) { w->sfb_lmax = 0; w->sfb_smax = 0; } else { w->sfb_lmax = 21; w->sfb_smax = 12; } /* gr_deco */
            memcpy(xr, out->xr[gr][ch], sizeof(xr));
            calc_xmin_ref(xr, out->ratio_l[gr][ch], out->ratio_s[gr][ch], w, sr, &xm);
            calc_scfsi_ref(L, xr, w, &xm, sr, ch, gr, out->scfsi[ch]);
            max_bits = resv_max_bits(L, out->pe[gr][ch], e->mean_bits, n_ch);
            out->max_bits[gr][ch] = max_bits;
            for (i = 0; i < 576; i++) { double a = fabs(xr[i]); if (a > mx) mx = a; }
            if (mx != 0.0) {
                w->q = (double)quantanf_init_ref(xr);
                outer_loop_ref(xr, max_bits, &xm, out->ix[gr][ch], w, out->scalefac_l[gr][ch], out->scalefac_s[gr][ch],
                               gr, sr, out->scfsi[ch], &wg[0][ch].g, out->scalefac_l[0][ch]);
            }
            L->resv_size += e->mean_bits / n_ch - w->g.part2_3_length;         /* ResvAdjust */
            w->g.global_gain = nint_ref(w->q + 210.0);
            L->addr[gr][ch][0] = w->g.address1; L->addr[gr][ch][1] = w->g.address2; L->addr[gr][ch][2] = w->g.address3;
        }
    { /* ResvFrameEnd, reservoir.c:155-226 */
        int over_bits, stuffing;
        if (n_ch == 2 && (e->mean_bits & 1)) L->resv_size += 1;
        over_bits = L->resv_size - L->resv_max;
        if (over_bits < 0) over_bits = 0;
        L->resv_size -= over_bits;
        stuffing = over_bits;
        if ((over_bits = L->resv_size % 8)) { stuffing += over_bits; L->resv_size -= over_bits; }
        if (stuffing) {
            if (wg[0][0].g.part2_3_length + stuffing < 4095) wg[0][0].g.part2_3_length += stuffing;
            else {
                for (gr = 0; gr < 2; gr++)
                    for (ch = 0; ch < n_ch; ch++) {
                        int extra, take;
                        if (stuffing == 0) break;
                        extra = 4095 - wg[gr][ch].g.part2_3_length;
                        take = extra < stuffing ? extra : stuffing;
                        wg[gr][ch].g.part2_3_length += take;
                        stuffing -= take;
                    }
                out->resv_drain = stuffing;
            }
        }
    }
    out->resv_size = L->resv_size;
    for (gr = 0; gr < 2; gr++)
        for (ch = 0; ch < n_ch; ch++) { out->gi[gr][ch] = wg[gr][ch].g; out->qstep[gr][ch] = wg[gr][ch].q; }
    return 0;
}

int l3o_encode_stream(int sfreq, int n_ch, int bitrate_kbps, const short *pcm, long n_samples, l3o_frame *out, long n_frames)
{
    l3o_enc *e = l3o_create(sfreq, n_ch, bitrate_kbps);
    short buf[2 * 1152];
    long f, i;
    int ch;
    if (!e) return -1;
    for (f = 0; f < n_frames; f++) {
        for (ch = 0; ch < n_ch; ch++)
            for (i = 0; i < 1152; i++) {
                long t = f * 1152 + i;
                buf[ch * 1152 + i] = (t < n_samples) ? pcm[ch * n_samples + t] : 0;
            }
        l3o_encode_frame(e, buf, &out[f]);
    }
    l3o_destroy(e);
    return 0;
}

/* ======================================================================================================
 * Bitstream formatter (TEST INFRASTRUCTURE): sequential restatement of III_format_bitstream
 * (l3bitstream.c:68-163), encodeSideInfo (:314-458), encodeMainData (:179-309), Huffmancodebits (:517-716),
 * HuffmanCode (:779-906), L3_huffman_coder_count1 (:728-767) and the frame assembler BF_BitstreamFrame /
 * WriteMainDataBits / BF_FlushBitstream (formatBitstream.c:53-125, 218-247).  Pinned against the byte streams
 * the unmodified reference CLI writes (tests/golden/cli_*.mp3).
 * ====================================================================================================== */
typedef struct {
    unsigned char *buf; long cap, nbits;          /* output byte stream */
    /* side-info queue (formatBitstream.c:276-398): each entry is the packed header+SI of one frame */
    unsigned char (*q)[40]; int q_head, q_tail, q_cap;
    int si_bits, frame_bits;
    long bit_count, this_frame, bits_remaining;   /* BitCount, ThisFrameSize, BitsRemaining */
} fmt_t;

static void fmt_put(fmt_t *F, unsigned val, int n) /* putbits, common.c:1010-1040: MSB first */
{
    int i;
    for (i = n - 1; i >= 0; i--) {
        if ((F->nbits >> 3) < F->cap && ((val >> i) & 1u)) F->buf[F->nbits >> 3] |= (unsigned char)(0x80 >> (F->nbits & 7));
        F->nbits++;
    }
}
static int fmt_write_side_info(fmt_t *F) /* formatBitstream.c:249-269 */
{
    int i;
    const unsigned char *e = F->q[F->q_head % F->q_cap];
    F->q_head++;
    F->this_frame = F->frame_bits;
    for (i = 0; i < F->si_bits / 8; i++) fmt_put(F, e[i], 8);
    return F->si_bits;
}
static void fmt_main(fmt_t *F, unsigned val, unsigned nbits) /* WriteMainDataBits, formatBitstream.c:218-247 */
{
    if (F->bit_count == F->this_frame) { F->bit_count = fmt_write_side_info(F); F->bits_remaining = F->this_frame - F->bit_count; }
    if (nbits == 0) return;
    if ((long)nbits > F->bits_remaining) {
        unsigned extra = val >> (nbits - F->bits_remaining);
        nbits -= (unsigned)F->bits_remaining;
        fmt_put(F, extra, (int)F->bits_remaining);
        F->bit_count = fmt_write_side_info(F);
        F->bits_remaining = F->this_frame - F->bit_count;
        fmt_put(F, val, (int)nbits);
    } else fmt_put(F, val, (int)nbits);
    F->bit_count += nbits; F->bits_remaining -= nbits;
}
/* bit packer for one side-info entry */
typedef struct { unsigned char *p; int n; } sip_t;
static void si_put(sip_t *s, unsigned v, int n) { int i; for (i = n - 1; i >= 0; i--) { if ((v >> i) & 1u) s->p[s->n >> 3] |= (unsigned char)(0x80 >> (s->n & 7)); s->n++; } }

static int fmt_huffman_pair(fmt_t *F, int t, int x, int y) /* HuffmanCode, l3bitstream.c:779-906 */
{
    unsigned signx = 0, signy = 0, code, ext = 0, idx, linbitsx = 0, linbitsy = 0;
    int cbits, xbits = 0;
    const mp3t_huff_desc *h = &MP3T_HUFF[t];
    if (t == 0) return 0;
    if (x < 0) { x = -x; signx = 1; }
    if (y < 0) { y = -y; signy = 1; }
    if (t > 15) {
        if (x > 14) { linbitsx = (unsigned)x - 15; x = 15; }
        if (y > 14) { linbitsy = (unsigned)y - 15; y = 15; }
        idx = (unsigned)x * h->ylen + (unsigned)y;
        code = MP3T_HCODE[h->off + idx]; cbits = MP3T_HLEN[h->off + idx];
        if (x > 14) { ext |= linbitsx; xbits += h->linbits; }
        if (x != 0) { ext <<= 1; ext |= signx; xbits += 1; }
        if (y > 14) { ext <<= h->linbits; ext |= linbitsy; xbits += h->linbits; }
        if (y != 0) { ext <<= 1; ext |= signy; xbits += 1; }
    } else {
        idx = (unsigned)x * h->ylen + (unsigned)y;
        code = MP3T_HCODE[h->off + idx]; cbits = MP3T_HLEN[h->off + idx];
        if (x != 0) { code <<= 1; code |= signx; cbits += 1; }
        if (y != 0) { code <<= 1; code |= signy; cbits += 1; }
    }
    if (cbits) fmt_main(F, code, (unsigned)cbits);
    if (xbits) fmt_main(F, ext, (unsigned)xbits);
    return cbits + xbits;
}

static int fmt_count1_quad(fmt_t *F, int t, int v, int w, int x, int y) /* l3bitstream.c:728-767 */
{
    unsigned sv = v < 0, sw = w < 0, sx = x < 0, sy = y < 0, p;
    int len, total;
    const mp3t_huff_desc *h = &MP3T_HUFF[t];
    v = abs(v); w = abs(w); x = abs(x); y = abs(y);
    p = (unsigned)(v + (w << 1) + (x << 2) + (y << 3));
    len = MP3T_HLEN[h->off + p];
    fmt_main(F, MP3T_HCODE[h->off + p], (unsigned)len);
    total = len;
    if (v) { fmt_main(F, sv, 1); total++; }
    if (w) { fmt_main(F, sw, 1); total++; }
    if (x) { fmt_main(F, sx, 1); total++; }
    if (y) { fmt_main(F, sy, 1); total++; }
    return total;
}

static void fmt_granule(fmt_t *F, const l3o_frame *fr, int gr, int ch, int sr)
{
    static const unsigned slen1_tab[16] = {0, 0, 0, 0, 3, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4};
    static const unsigned slen2_tab[16] = {0, 1, 2, 3, 0, 1, 2, 3, 1, 2, 3, 1, 2, 3, 2, 3};
    const l3o_gr_info *g = &fr->gi[gr][ch];
    unsigned slen1 = slen1_tab[g->scalefac_compress], slen2 = slen2_tab[g->scalefac_compress];
    int ix[576], i, sfb, w, bits = 0, bigvalues, count1_end, stuffing;
    for (i = 0; i < 576; i++) { ix[i] = fr->ix[gr][ch][i]; if (fr->xr[gr][ch][i] < 0 && ix[i] > 0) ix[i] = -ix[i]; } /* :115-125 */
    /* scalefactors, l3bitstream.c:196-251 (no mixed blocks) */
    if (g->window_switching_flag == 1 && g->block_type == 2) {
        for (sfb = 0; sfb < 6; sfb++) for (w = 0; w < 3; w++) fmt_main(F, (unsigned)fr->scalefac_s[gr][ch][sfb][w], slen1);
        for (sfb = 6; sfb < 12; sfb++) for (w = 0; w < 3; w++) fmt_main(F, (unsigned)fr->scalefac_s[gr][ch][sfb][w], slen2);
    } else {
        const int *sc = fr->scfsi[ch];
        if (gr == 0 || sc[0] == 0) for (sfb = 0; sfb < 6; sfb++) fmt_main(F, (unsigned)fr->scalefac_l[gr][ch][sfb], slen1);
        if (gr == 0 || sc[1] == 0) for (sfb = 6; sfb < 11; sfb++) fmt_main(F, (unsigned)fr->scalefac_l[gr][ch][sfb], slen1);
        if (gr == 0 || sc[2] == 0) for (sfb = 11; sfb < 16; sfb++) fmt_main(F, (unsigned)fr->scalefac_l[gr][ch][sfb], slen2);
        if (gr == 0 || sc[3] == 0) for (sfb = 16; sfb < 21; sfb++) fmt_main(F, (unsigned)fr->scalefac_l[gr][ch][sfb], slen2);
    }
    /* Huffmancodebits, l3bitstream.c:517-716 */
    bigvalues = g->big_values * 2;
    if (bigvalues) {
        if (g->window_switching_flag && g->block_type == 2) {
            const short *sf = MP3T_SFB_SHORT[sr];
            for (sfb = 0; sfb < 13; sfb++) {
                int start = sf[sfb], end = sf[sfb + 1], line;
                int t = (start < 12) ? g->table_select[0] : g->table_select[1];
                for (w = 0; w < 3; w++)
                    for (line = start; line < end; line += 2) bits += fmt_huffman_pair(F, t, ix[line * 3 + w], ix[(line + 1) * 3 + w]);
            }
        } else {
            const short *sf = MP3T_SFB_LONG[sr];
            int r1 = sf[g->region0_count + 1], r2 = sf[g->region0_count + 1 + g->region1_count + 1];
            for (i = 0; i < bigvalues; i += 2) {
                int t = (i < r1) ? g->table_select[0] : (i < r2) ? g->table_select[1] : g->table_select[2];
                if (t) bits += fmt_huffman_pair(F, t, ix[i], ix[i + 1]);
            }
        }
    }
    count1_end = bigvalues + g->count1 * 4;
    for (i = bigvalues; i < count1_end; i += 4) bits += fmt_count1_quad(F, 32 + g->count1table_select, ix[i], ix[i + 1], ix[i + 2], ix[i + 3]);
    stuffing = g->part2_3_length - g->part2_length - bits;
    if (stuffing > 0) { while (stuffing >= 32) { fmt_main(F, ~0u, 32); stuffing -= 32; } if (stuffing) fmt_main(F, ~0u, (unsigned)stuffing); }
}

/* Whole stream: frames (from l3o_encode_stream) -> the byte stream the reference CLI writes, INCLUDING the trailing
 * byte close_bit_stream_w() emits (common.c:968-974 writes buf_byte_idx+1 bytes).  main_data_begin of every frame
 * is returned in mdb[n_frames] when non-NULL.  Returns the number of bytes, or -1 if `cap` is too small. */
long l3o_format_stream(int sfreq, int n_ch, int bitrate_kbps, const l3o_frame *frames, long n_frames, unsigned char *out, long cap,
                       int *mdb)
{
    static const int rates[15] = {0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320};
    fmt_t F;
    long f, len;
    int gr, ch, i, br_idx = 0, sr = (sfreq == 32000) ? 0 : (sfreq == 44100) ? 1 : 2;
    int sf_idx = (sfreq == 44100) ? 0 : (sfreq == 48000) ? 1 : 2;   /* header index, common.c s_freq[] order */
    int main_data_begin = 0;
    init_common();
    for (i = 1; i < 15; i++) if (rates[i] == bitrate_kbps) br_idx = i;
    memset(&F, 0, sizeof(F));
    F.buf = out; F.cap = cap; memset(out, 0, (size_t)cap);
    F.q_cap = 64; F.q = calloc((size_t)F.q_cap, 40);
    F.si_bits = 32 + (n_ch == 2 ? 256 : 136);
    F.frame_bits = 8 * (int)((1152.0 / (sfreq / 1000.0)) * ((double)bitrate_kbps / 8.0));
    for (f = 0; f < n_frames; f++) {
        const l3o_frame *fr = &frames[f];
        sip_t s;
        int elements = 0, fwd_frame = 0, fwd_si = 0;
        /* encodeSideInfo + store_side_info */
        if (F.q_tail - F.q_head >= F.q_cap) { free(F.q); return -1; }
        s.p = F.q[F.q_tail % F.q_cap]; s.n = 0; memset(s.p, 0, 40);
        si_put(&s, 0xfff, 12); si_put(&s, 1, 1); si_put(&s, 4 - 3, 2); si_put(&s, 1, 1); si_put(&s, (unsigned)br_idx, 4);
        si_put(&s, (unsigned)sf_idx, 2); si_put(&s, 0, 1); si_put(&s, 0, 1); si_put(&s, n_ch == 2 ? 0 : 3, 2); si_put(&s, 0, 2);
        si_put(&s, 0, 1); si_put(&s, 0, 1); si_put(&s, 0, 2);
        si_put(&s, (unsigned)main_data_begin, 9); si_put(&s, 0, n_ch == 2 ? 3 : 5);
        if (mdb) mdb[f] = main_data_begin;
        for (ch = 0; ch < n_ch; ch++) for (i = 0; i < 4; i++) si_put(&s, (unsigned)fr->scfsi[ch][i], 1);
        for (gr = 0; gr < 2; gr++)
            for (ch = 0; ch < n_ch; ch++) {
                const l3o_gr_info *g = &fr->gi[gr][ch];
                si_put(&s, (unsigned)g->part2_3_length, 12); si_put(&s, (unsigned)g->big_values, 9); si_put(&s, (unsigned)g->global_gain, 8);
                si_put(&s, (unsigned)g->scalefac_compress, 4); si_put(&s, (unsigned)g->window_switching_flag, 1);
                if (g->window_switching_flag) {
                    si_put(&s, (unsigned)g->block_type, 2); si_put(&s, (unsigned)g->mixed_block_flag, 1);
                    si_put(&s, (unsigned)g->table_select[0], 5); si_put(&s, (unsigned)g->table_select[1], 5);
                    si_put(&s, 0, 3); si_put(&s, 0, 3); si_put(&s, 0, 3);      /* subblock_gain, always 0 */
                } else {
                    si_put(&s, (unsigned)g->table_select[0], 5); si_put(&s, (unsigned)g->table_select[1], 5); si_put(&s, (unsigned)g->table_select[2], 5);
                    si_put(&s, (unsigned)g->region0_count, 4); si_put(&s, (unsigned)g->region1_count, 3);
                }
                si_put(&s, (unsigned)g->preflag, 1); si_put(&s, (unsigned)g->scalefac_scale, 1); si_put(&s, (unsigned)g->count1table_select, 1);
            }
        F.q_tail++;
        /* main_data(), formatBitstream.c:187-205 */
        for (gr = 0; gr < 2; gr++) for (ch = 0; ch < n_ch; ch++) fmt_granule(&F, fr, gr, ch, sr);
        { int d = fr->resv_drain; while (d >= 32) { fmt_main(&F, 0, 32); d -= 32; } if (d) fmt_main(&F, 0, (unsigned)d); }   /* :497-513 */
        /* nextBackPtr, formatBitstream.c:76-79 */
        for (i = F.q_head; i < F.q_tail; i++) { elements++; fwd_frame += F.frame_bits; fwd_si += F.si_bits; }
        main_data_begin = (int)(F.bits_remaining / 8) + fwd_frame / 8 - fwd_si / 8;
    }
    /* BF_FlushBitstream, formatBitstream.c:87-125 */
    if (F.q_tail > F.q_head) {
        long rem = (long)(F.q_tail - F.q_head) * (F.frame_bits - F.si_bits);
        while (rem >= 32) { fmt_main(&F, 0, 32); rem -= 32; }
        fmt_main(&F, 0, (unsigned)rem);
    }
    free(F.q);
    len = (F.nbits >> 3) + 1;   /* empty_buffer(bs, buf_byte_idx) writes the partially filled byte too */
    return len <= cap ? len : -1;
}

/* ---- table access for the Python test decoder (oracle/mp3dec.py) ---- */
int l3o_huff_table(int t, int *xlen, int *ylen, int *linbits, unsigned *codes, unsigned char *lens)
{
    int i, n;
    if (t < 0 || t > 33) return -1;
    *xlen = MP3T_HUFF[t].xlen; *ylen = MP3T_HUFF[t].ylen; *linbits = MP3T_HUFF[t].linbits;
    n = (t >= 32) ? 16 : MP3T_HUFF[t].xlen * MP3T_HUFF[t].ylen;
    for (i = 0; i < n; i++) { codes[i] = MP3T_HCODE[MP3T_HUFF[t].off + i]; lens[i] = MP3T_HLEN[MP3T_HUFF[t].off + i]; }
    return n;
}
void l3o_sfb_tables(int sr_idx, short *l23, short *s14)
{
    int i;
    for (i = 0; i < 23; i++) l23[i] = MP3T_SFB_LONG[sr_idx][i];
    for (i = 0; i < 14; i++) s14[i] = MP3T_SFB_SHORT[sr_idx][i];
}
void l3o_analysis_window(double *c512) { int i; for (i = 0; i < 512; i++) c512[i] = MP3T_ANA_WINDOW[i]; }
