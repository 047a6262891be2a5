// front_core.h — polyphase analysis filterbank + MDCT + alias reduction, fused, FP64 exact path.
//
// Replaces window_subband()/filter_subband() (/root/reference/src/encode.c:287-409) and
// mdct_sub()/mdct() (mdct.c:25-198) for the batched entry points.  One warp walks T consecutive
// granules of one (stream, channel).  Lane == subband == MDCT band, so the 32 subband samples of
// a time slot are produced one per lane and the 36 MDCT inputs of a band never leave their lane:
// subband samples are NOT written to HBM unless the caller asks for them (parity tests).
//
// Arithmetic follows the reference operation for operation (unfused mul/add, same summation
// order) so subband samples are bit-identical; MDCT uses the plain dot product for every long
// block type (the reference's hand-unrolled type-0 form differs only in summation order, <=2e-14).
#pragma once
#include "simt.h"
#include "tables.h"

namespace mp3gpu {

using simt::PerThread;
using simt::WarpCtx;

struct FrontWarpSmem {
    double ring[512];   // scaled PCM, sample t at ring[t & 511]
    double y[64];
    double ys[32];      // ysum[0..15], ysub[0..14] at [16..30]
    double xr[576];     // staging for alias reduction + coalesced store
};

// one polyphase slot: 32 new samples (time t0..t0+31) -> s (one subband sample per lane)
SIMT_FN void polyphase_slot(const WarpCtx &w, const double *window, FrontWarpSmem &M, const PerThread<double> am[31],
                            const short *pcm32, long t0, PerThread<double> &s_out)
{
    FOR_THREADS(w)
    M.ring[(t0 + lane) & 511] = (double)pcm32[lane] / 32768;                       // encode.c:306-307
    END_THREADS
    w.sync();
    const long tnew = t0 + 31;
    FOR_THREADS(w)
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int i = lane + 32 * h;
        // z[i+64j] = x * enwindow (encode.c:310-311); y[i] = z[i] + z[i+64] + ... (encode.c:392-396)
        double acc = simt::dmul(M.ring[(tnew - i) & 511], window[i]);
#pragma unroll
        for (int j = 1; j < 8; j++)
            acc = simt::dadd(acc, simt::dmul(M.ring[(tnew - i - 64 * j) & 511], window[i + 64 * j]));
        M.y[i] = acc;
    }
    END_THREADS
    w.sync();
    FOR_THREADS(w)
    if (lane < 16) M.ys[lane] = simt::dadd(M.y[lane], M.y[32 - lane]);              // encode.c:397
    else if (lane < 31) M.ys[lane] = simt::dsub(M.y[33 + lane - 16], M.y[63 - (lane - 16)]);  // encode.c:398
    END_THREADS
    w.sync();
    FOR_THREADS(w)
    double si = M.y[16];                                                            // encode.c:399-408
#pragma unroll
    for (int j = 0; j < 31; j++) si = simt::dadd(si, simt::dmul(am[j](), M.ys[j]));
    s_out() = si;
    END_THREADS
    w.sync();
}

// MDCT of one band held in registers: in[0..17] = previous granule, in[18..35] = current (sign-fixed)
SIMT_FN void mdct_lane(const FrontTables &F, const double in[36], int bt, double out[18])
{
    if (bt == 2) {                                                                  // mdct.c:171-185
#pragma unroll
        for (int l = 0; l < 3; l++)
#pragma unroll
            for (int m = 0; m < 6; m++) {
                double sum = 0.0;
#pragma unroll
                for (int k = 0; k < 12; k++)
                    sum = simt::dadd(sum, simt::dmul(simt::dmul(F.win[2][k], in[k + 6 * l + 6]), F.cos_s[m][k]));
                out[3 * m + l] = sum;
            }
    } else {                                                                        // mdct.c:188-198
        double fin[36];
#pragma unroll
        for (int k = 0; k < 36; k++) fin[k] = simt::dmul(F.win[bt][k], in[k]);
#pragma unroll
        for (int m = 0; m < 18; m++) {
            double sum = 0.0;
#pragma unroll
            for (int k = 0; k < 36; k++) sum = simt::dadd(sum, simt::dmul(fin[k], F.cos_l[m][k]));
            out[m] = sum;
        }
    }
}

// MDCT of all 32 bands (lane == band) + alias reduction + coalesced store; prev <- cur afterwards.
SIMT_FN void mdct_store(const WarpCtx &w, const FrontTables &F, FrontWarpSmem &M, PerThread<double> prev[18],
                        const PerThread<double> cur[18], int bt, double *xr_out)
{
    FOR_THREADS(w)
    double in[36], out[18];
#pragma unroll
    for (int k = 0; k < 18; k++) { in[k] = prev[k](); in[k + 18] = cur[k](); }
    mdct_lane(F, in, bt, out);
#pragma unroll
    for (int m = 0; m < 18; m++) M.xr[lane * 18 + m] = out[m];
#pragma unroll
    for (int k = 0; k < 18; k++) prev[k]() = cur[k]();
    END_THREADS
    w.sync();
    if (bt != 2) {                                                                  // mdct.c:83-91
        FOR_THREADS(w)
        if (lane < 31) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                double a = M.xr[lane * 18 + 17 - k], b = M.xr[(lane + 1) * 18 + k];
                double bu = simt::dadd(simt::dmul(a, F.cs[k]), simt::dmul(b, F.ca[k]));
                double bd = simt::dsub(simt::dmul(b, F.cs[k]), simt::dmul(a, F.ca[k]));
                M.xr[lane * 18 + 17 - k] = bu;
                M.xr[(lane + 1) * 18 + k] = bd;
            }
        }
        END_THREADS
        w.sync();
    }
    FOR_THREADS(w)
    for (int i = lane; i < 576; i += 32) xr_out[i] = M.xr[i];
    END_THREADS
    w.sync();
}

// Walk granules [g_first, g_first + n_gran) of one channel.
//  pcm        : channel samples; pcm[t] valid for t >= -hist (zeros before the stream started)
//  block_type : per granule (stride bt_stride ints), produced by the psy scan
//  xr         : output, 576 doubles per granule (stride xr_stride doubles)
//  sb_out     : optional raw subband samples [granule][18][32] (parity tests), stride sb_stride
// window: the 512 analysis-window taps (shared memory copy on the device: lane-varying index)
SIMT_FN void front_walk(const WarpCtx &w, const FrontTables &F, const double *window, FrontWarpSmem &M, const short *pcm, long g_first, int n_gran,
                        const int *block_type, long bt_stride, double *xr, long xr_stride, double *sb_out, long sb_stride,
                        bool do_mdct = true)
{
    PerThread<double> am[31];
    FOR_THREADS(w)
#pragma unroll
    for (int j = 0; j < 31; j++) am[j]() = F.am[lane][j];
    END_THREADS
    // ring warm-up: the 480 samples before the first slot we compute (granule g_first-1, slot 0)
    const long t_begin = 576 * (g_first - 1);
    FOR_THREADS(w)
    for (int i = lane; i < 480; i += 32) {
        long t = t_begin - 480 + i;
        M.ring[t & 511] = (double)pcm[t] / 32768;
    }
    END_THREADS
    w.sync();
    PerThread<double> prev[18], cur[18];
    // previous granule's subband samples (mdct.c:68-72 reads slot gr, saved at :99-102)
    for (int k = 0; k < 18; k++) {
        PerThread<double> s;
        if (!do_mdct) { FOR_THREADS(w) M.ring[(t_begin + 32 * k + lane) & 511] = (double)pcm[t_begin + 32 * k + lane] / 32768; END_THREADS w.sync(); continue; }
        polyphase_slot(w, window, M, am, pcm + t_begin + 32 * k, t_begin + 32 * k, s);
        FOR_THREADS(w)
        prev[k]() = ((lane & 1) && (k & 1)) ? simt::dmul(s(), -1.0) : s();          // mdct.c:57-60
        END_THREADS
    }
    for (int g = 0; g < n_gran; g++) {
        const long t0 = 576 * (g_first + g);
        for (int k = 0; k < 18; k++) {
            PerThread<double> s;
            polyphase_slot(w, window, M, am, pcm + t0 + 32 * k, t0 + 32 * k, s);
            FOR_THREADS(w)
            if (sb_out) sb_out[g * sb_stride + k * 32 + lane] = s();
            cur[k]() = ((lane & 1) && (k & 1)) ? simt::dmul(s(), -1.0) : s();
            END_THREADS
        }
        if (do_mdct) mdct_store(w, F, M, prev, cur, block_type[g * bt_stride], xr + g * xr_stride);
    }
}

}  // namespace mp3gpu
