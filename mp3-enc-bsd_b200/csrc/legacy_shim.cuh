// legacy_shim.cuh — the reference's single-frame entry points (include/mp3gpu_legacy.h) on top of the same
// kernels as the batched API.  Included by mp3gpu.cu only.  Every function cites the reference lines whose
// caller-visible behaviour it reproduces; the arithmetic itself runs on the device.
#pragma once
#include "../../include/mp3gpu_legacy.h"

namespace mp3gpu {

// ---- window_subband / filter_subband, one slot per call (encode.c:287-316, 361-409) ----------------
__global__ void k_legacy_window(double *ring /*[512]*/, int off, const short *pcm32, double *z)
{
    __shared__ double x[512];
    const int t = threadIdx.x;                      // 512 threads
    x[t] = ring[t];
    __syncthreads();
    if (t < 32) x[31 - t + off] = (double)pcm32[t] / 32768;                  // encode.c:306-307
    __syncthreads();
    z[t] = __dmul_rn(x[(t + off) & 511], c_front.window[t]);                  // encode.c:310-311
    ring[t] = x[t];
}

__global__ void k_legacy_filter(const double *z, double *s)
{
    __shared__ double y[64], ys[32];
    const int t = threadIdx.x;                      // 64 threads
    double acc = z[t];                                                        // encode.c:392-396
#pragma unroll
    for (int j = 1; j < 8; j++) acc = __dadd_rn(acc, z[t + 64 * j]);
    y[t] = acc;
    __syncthreads();
    if (t < 16) ys[t] = __dadd_rn(y[t], y[32 - t]);                           // encode.c:397
    else if (t < 31) ys[t] = __dsub_rn(y[33 + t - 16], y[63 - (t - 16)]);     // encode.c:398
    __syncthreads();
    if (t < 32) {
        double si = y[16];                                                    // encode.c:399-408
        for (int j = 0; j < 31; j++) si = __dadd_rn(si, __dmul_rn(c_front.am[t][j], ys[j]));
        s[t] = si;
    }
}

// ---- mdct_sub for one frame: sb [ch][3][18][32] already sign-fixed, xr [gr][ch][576] ----------------
__global__ void __launch_bounds__(32)
k_legacy_mdct(const double *sb, const int *block_type /*[gr*2+ch]*/, int stereo, double *xr)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrontWarpSmem &M = *reinterpret_cast<FrontWarpSmem *>(smem_raw);
    const int gr = blockIdx.x / stereo, ch = blockIdx.x % stereo, lane = threadIdx.x;
    WarpCtx w;
    PerThread<double> prev[18], cur[18];
#pragma unroll
    for (int k = 0; k < 18; k++) {
        prev[k].v = sb[((ch * 3 + gr) * 18 + k) * 32 + lane];
        cur[k].v = sb[((ch * 3 + gr + 1) * 18 + k) * 32 + lane];
    }
    mdct_store(w, c_front, M, prev, cur, block_type[gr * 2 + ch], xr + (gr * 2 + ch) * 576);
}

struct LegacyState {
    bool ready = false;
    long launches = 0;
    // filterbank
    double *d_ring = nullptr;   // [2][512]
    int off[2] = {0, 0};
    short *d_pcm32 = nullptr;
    double *d_z = nullptr, *d_s = nullptr;
    // mdct
    double *d_sb = nullptr, *d_xr = nullptr;
    int *d_bt = nullptr;
    // psy
    int psy_sr = -1;
    PsyTables *d_psy_tab = nullptr;
    uint32_t *ops1024 = nullptr, *ops256 = nullptr;
    int *lv1024 = nullptr, *lv256 = nullptr;
    uint32_t *out1024 = nullptr, *out256 = nullptr;
    FftTwiddle *d_tw = nullptr;
    PsyDev psy_dev;
    short *d_save = nullptr;            // [2][1344]
    PsyMid *d_mid = nullptr;
    PsyChanState *d_psy_state = nullptr;  // [2]
    PsyOut *d_psyout = nullptr;           // [4]
    // rate loop
    int loop_sr = -1;
    RateTables *d_rate_tab = nullptr;
    LoopStreamState *d_loop_state = nullptr;
    LoopLaneState *d_lane_state = nullptr;
    short *d_ix = nullptr;
    GrInfoOut *d_gi = nullptr;
    unsigned char *d_sf = nullptr;
    FrameOut *d_fo = nullptr;
    int *d_sched = nullptr;     // work queue of the persistent rate loop (one stream: ticket counter + one progress word)
    double *d_xr4 = nullptr;
};
static LegacyState g_legacy;
static unsigned g_legacy_partition_table[4] = {0, 0, 0, 0};

// The legacy entry points return void and the reference's host program cannot check a status, so they keep
// the reference's own convention for unrecoverable conditions: message + exit(1) (musicin.c:550-557,
// l3psy.c:174-175).  There is no CPU fallback to fall back to.
[[noreturn]] static void legacy_fatal(const char *msg)
{
    fprintf(stderr, "libmp3gpu (legacy entry point): %s\n", msg);
    exit(1);
}
static int legacy_fail(int code, const char *what, cudaError_t e)
{
    fail(code, "legacy %s: %s", what, cudaGetErrorString(e));
    legacy_fatal(g_err);
    return code;
}
#define LCU(call)                                                                  \
    do {                                                                           \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess) { legacy_fail(MP3GPU_ECUDA, #call, e_); return; }   \
    } while (0)

static bool legacy_init()
{
    LegacyState &L = g_legacy;
    if (L.ready) return true;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) legacy_fatal("no CUDA device: libmp3gpu has no CPU fallback");
    FrontTables *F = new FrontTables;
    build_front_tables(F);
    cudaError_t e = cudaMemcpyToSymbol(c_front, F, sizeof(FrontTables));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_front, F, sizeof(FrontTables));
    delete F;
    if (e == cudaSuccess) e = cudaMalloc(&L.d_ring, 2 * 512 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&L.d_pcm32, 32 * sizeof(short));
    if (e == cudaSuccess) e = cudaMalloc(&L.d_z, 512 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&L.d_s, 32 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&L.d_sb, 2 * 3 * 576 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&L.d_xr, 4 * 576 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&L.d_bt, 4 * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&L.d_save, 2 * 1344 * sizeof(short));
    if (e == cudaSuccess) e = cudaMalloc(&L.d_mid, sizeof(PsyMid));
    if (e == cudaSuccess) e = cudaMalloc(&L.d_psy_state, 2 * sizeof(PsyChanState));
    if (e == cudaSuccess) e = cudaMalloc(&L.d_psyout, 4 * sizeof(PsyOut));
    if (e == cudaSuccess) e = cudaMalloc(&L.d_loop_state, sizeof(LoopStreamState));
    if (e == cudaSuccess) e = cudaMalloc(&L.d_lane_state, sizeof(LoopLaneState));
    if (e == cudaSuccess) e = cudaMalloc(&L.d_ix, 4 * 576 * sizeof(short));
    if (e == cudaSuccess) e = cudaMalloc(&L.d_gi, 4 * sizeof(GrInfoOut));
    if (e == cudaSuccess) e = cudaMalloc(&L.d_sf, 4 * 40);
    if (e == cudaSuccess) e = cudaMalloc(&L.d_fo, sizeof(FrameOut));
    if (e == cudaSuccess) e = cudaMalloc(&L.d_sched, 2 * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&L.d_xr4, 4 * 576 * sizeof(double));
    if (e != cudaSuccess) { legacy_fail(MP3GPU_ECUDA, "init", e); return false; }
    cudaFuncSetAttribute(k_psy_front, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(PSYF_WARPS * sizeof(PsyFrontSmem)));
    cudaFuncSetAttribute(k_rate_loop<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RL_SMEM_BYTES);
    L.ready = true;
    mp3gpu_legacy_reset();
    return true;
}

}  // namespace mp3gpu

using namespace mp3gpu;

extern "C" long mp3gpu_legacy_kernel_launches(void) { return g_legacy.launches; }

extern "C" void mp3gpu_legacy_reset(void)
{
    LegacyState &L = g_legacy;
    if (!L.ready) return;
    L.off[0] = L.off[1] = 0;
    cudaMemset(L.d_ring, 0, 2 * 512 * sizeof(double));
    cudaMemset(L.d_psy_state, 0, 2 * sizeof(PsyChanState));
    cudaMemset(L.d_psyout, 0, 4 * sizeof(PsyOut));
    cudaMemset(L.d_loop_state, 0, sizeof(LoopStreamState));
    cudaMemset(L.d_lane_state, 0, sizeof(LoopLaneState));
}

// encode.c:287-316: consumes 32 samples through the caller's pointer (which it advances), updates the
// per-channel ring, returns the windowed vector z.
extern "C" void window_subband(short **buffer, double z[512], int k)
{
    if (!legacy_init()) return;
    LegacyState &L = g_legacy;
    LCU(cudaMemcpy(L.d_pcm32, *buffer, 32 * sizeof(short), cudaMemcpyHostToDevice));
    *buffer += 32;
    k_legacy_window<<<1, 512>>>(L.d_ring + 512 * k, L.off[k], L.d_pcm32, L.d_z);
    L.launches++;
    LCU(cudaGetLastError());
    LCU(cudaMemcpy(z, L.d_z, 512 * sizeof(double), cudaMemcpyDeviceToHost));
    L.off[k] = (L.off[k] + 480) & 511;                                        // encode.c:313-314
}

// encode.c:361-409
extern "C" void filter_subband(double z[512], double s[32])
{
    if (!legacy_init()) return;
    LegacyState &L = g_legacy;
    LCU(cudaMemcpy(L.d_z, z, 512 * sizeof(double), cudaMemcpyHostToDevice));
    k_legacy_filter<<<1, 64>>>(L.d_z, L.d_s);
    L.launches++;
    LCU(cudaGetLastError());
    LCU(cudaMemcpy(s, L.d_s, 32 * sizeof(double), cudaMemcpyDeviceToHost));
}

// mdct.c:25-103: sign fix in place (:57-60), MDCT + alias per (gr, ch), slot mode_gr saved to slot 0 (:99-102)
extern "C" void mdct_sub(mp3gpu_L3SBS *sb_sample, double (*mdct_freq)[2][576], int stereo, III_side_info_t *l3_side, int mode_gr)
{
    if (!legacy_init()) return;
    LegacyState &L = g_legacy;
    int bt[4] = {0, 0, 0, 0};
    for (int gr = 0; gr < mode_gr; gr++)
        for (int ch = 0; ch < stereo; ch++) {
            bt[gr * 2 + ch] = (int)l3_side->gr[gr].ch[ch].tt.block_type;
            for (int band = 1; band < 32; band += 2)
                for (int k = 1; k < 18; k += 2) (*sb_sample)[ch][gr + 1][k][band] *= -1.0;
        }
    LCU(cudaMemcpy(L.d_sb, sb_sample, sizeof(mp3gpu_L3SBS), cudaMemcpyHostToDevice));
    LCU(cudaMemcpy(L.d_bt, bt, sizeof(bt), cudaMemcpyHostToDevice));
    k_legacy_mdct<<<mode_gr * stereo, 32, sizeof(FrontWarpSmem)>>>(L.d_sb, L.d_bt, stereo, L.d_xr);
    L.launches++;
    LCU(cudaGetLastError());
    double xr[4][576];
    LCU(cudaMemcpy(xr, L.d_xr, sizeof(xr), cudaMemcpyDeviceToHost));
    for (int gr = 0; gr < mode_gr; gr++)
        for (int ch = 0; ch < stereo; ch++) memcpy(mdct_freq[gr][ch], xr[gr * 2 + ch], 576 * sizeof(double));
    for (int ch = 0; ch < stereo; ch++) memcpy((*sb_sample)[ch][0], (*sb_sample)[ch][mode_gr], 576 * sizeof(double));
}

// l3psy.c:53-764 (Layer III branch :443-740): one granule of one channel.  `savebuf` is the caller's delay
// line and is updated exactly as the reference does (:477-481).
extern "C" void L3psycho_anal(short *buffer, short savebuf[1344], int chn, int lay, float snr32[32], double sfreq,
                              double ratio_d[21], double ratio_ds[12][3], double *pe, gr_info *cod_info)
{
    (void)snr32;
    if (!legacy_init()) return;
    LegacyState &L = g_legacy;
    if (lay != 3) legacy_fatal("L3psycho_anal: only Layer III is implemented");
    const int sr = sr_index((int)(sfreq + 0.5));
    if (sr < 0) legacy_fatal("L3psycho_anal: invalid sampling frequency (MPEG-1 rates only)");   // l3psy.c:169-176 exit()s too
    if (L.psy_sr != sr) {
        if (L.psy_sr >= 0) legacy_fatal("L3psycho_anal: sampling frequency changed between calls");
        PsyTables *P = new PsyTables;
        build_psy_tables(sr, P);
        cudaError_t e = cudaMalloc(&L.d_psy_tab, sizeof(PsyTables));
        if (e == cudaSuccess) e = cudaMemcpy(L.d_psy_tab, P, sizeof(PsyTables), cudaMemcpyHostToDevice);
        delete P;
        std::vector<FftTwiddle> tw; std::vector<int> base;
        build_fft_twiddles(&tw, &base);
        FftProgram P10, P8;
        build_fft_program(10, base, &P10);
        build_fft_program(8, base, &P8);
        if (e == cudaSuccess) e = cudaMalloc(&L.d_tw, tw.size() * sizeof(FftTwiddle));
        if (e == cudaSuccess) e = cudaMemcpy(L.d_tw, tw.data(), tw.size() * sizeof(FftTwiddle), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { legacy_fail(MP3GPU_ECUDA, "psy tables", e); return; }
        if (upload_fft(P10, &L.ops1024, &L.lv1024, &L.out1024, &L.psy_dev.f1024) ||
            upload_fft(P8, &L.ops256, &L.lv256, &L.out256, &L.psy_dev.f256)) legacy_fatal(g_err);
        L.psy_dev.T = L.d_psy_tab; L.psy_dev.tw = L.d_tw;
        L.psy_sr = sr;
    }
    memmove(savebuf, savebuf + 576, 768 * sizeof(short));                    // l3psy.c:477-478
    memcpy(savebuf + 768, buffer, 576 * sizeof(short));                      // l3psy.c:480-481
    short *row = L.d_save + 1344 * chn;
    LCU(cudaMemcpy(row, savebuf, 1344 * sizeof(short), cudaMemcpyHostToDevice));
    // the granule's first new sample is savebuf[768]; psy_front reads [-768, 576) around it
    k_psy_front<<<1, PSYF_WARPS * 32, PSYF_WARPS * sizeof(PsyFrontSmem)>>>(L.psy_dev, row + 768 - HIST, 0, 0, 1, 1, 1, nullptr, L.d_mid);
    k_psy_scan<<<1, PSYS_WARPS * 32, PSYS_WARPS * sizeof(PsyScanSmem)>>>(L.d_psy_tab, L.d_mid, L.d_psy_state + chn, 1, 1, 1, nullptr, L.d_psyout);
    L.launches += 2;
    LCU(cudaGetLastError());
    PsyOut po;
    LCU(cudaMemcpy(&po, L.d_psyout, sizeof(po), cudaMemcpyDeviceToHost));
    for (int j = 0; j < 21; j++) ratio_d[j] = po.ratio_l[j];                 // l3psy.c:452-456 (previous call's ratios)
    for (int j = 0; j < 12; j++)
        for (int i = 0; i < 3; i++) ratio_ds[j][i] = po.ratio_s[3 * j + i];
    *pe = po.pe;
    cod_info->block_type = (unsigned)po.block_type;                          // l3psy.c:732-739
    cod_info->window_switching_flag = po.block_type != 0;
    cod_info->mixed_block_flag = 0;
}

// loop.c:232-362 for one frame; the reservoir recurrence (reservoir.c) runs inside the kernel.
extern "C" void iteration_loop(double pe[][2], double xr_org[2][2][576], III_psy_ratio *ratio, III_side_info_t *l3_side,
                               int l3_enc[2][2][576], int mean_bits, int stereo, double xr_dec[2][2][576],
                               III_scalefac_t *scalefac, frame_params *fr_ps, int ancillary_pad, int bitsPerFrame)
{
    (void)xr_dec; (void)ancillary_pad;
    if (!legacy_init()) return;
    LegacyState &L = g_legacy;
    const layer *info = fr_ps->header;
    if (info->version != 1) legacy_fatal("iteration_loop: MPEG-1 only");
    static const int idx2sr[3] = {1, 2, 0};   // header index 0:44.1 1:48 2:32 kHz -> table index 0:32 1:44.1 2:48
    const int sr = idx2sr[info->sampling_frequency % 3];
    if (L.loop_sr != sr) {
        RateTables *R = new RateTables;
        build_rate_tables(sr, R);
        cudaError_t e = L.d_rate_tab ? cudaSuccess : cudaMalloc(&L.d_rate_tab, sizeof(RateTables));
        if (e == cudaSuccess) e = cudaMemcpy(L.d_rate_tab, R, sizeof(RateTables), cudaMemcpyHostToDevice);
        delete R;
        if (e != cudaSuccess) { legacy_fail(MP3GPU_ECUDA, "rate tables", e); return; }
        L.loop_sr = sr;
    }
    PsyOut po[4];
    double xr[4][576];
    memset(po, 0, sizeof(po));
    for (int gr = 0; gr < 2; gr++)
        for (int ch = 0; ch < stereo; ch++) {
            PsyOut &p = po[gr * stereo + ch];
            p.pe = pe[gr][ch];
            memcpy(p.ratio_l, ratio->l[gr][ch], 21 * sizeof(double));
            memcpy(p.ratio_s, ratio->s[gr][ch], 36 * sizeof(double));
            p.block_type = (int)l3_side->gr[gr].ch[ch].tt.block_type;
            memcpy(xr[gr * stereo + ch], xr_org[gr][ch], 576 * sizeof(double));
        }
    LCU(cudaMemcpy(L.d_psyout, po, sizeof(po), cudaMemcpyHostToDevice));
    LCU(cudaMemcpy(L.d_xr4, xr, sizeof(xr), cudaMemcpyHostToDevice));
    FrameGeom G;
    G.n_ch = stereo; G.mean_bits = mean_bits; G.bits_per_frame = bitsPerFrame;
    frame_geom_derive(&G);
    LCU(cudaMemset(L.d_sched, 0, 2 * sizeof(int)));
    SegArgs no_seg;
    memset(&no_seg, 0, sizeof(no_seg));
    k_rate_loop<false><<<1, 32, RL_HOT_BYTES + sizeof(RateWarpSmem)>>>(L.d_rate_tab, G, L.d_loop_state, L.d_lane_state, 1, 1, nullptr, L.d_sched, no_seg,
                                                                 L.d_xr4, L.d_psyout, L.d_ix, L.d_gi, L.d_sf, L.d_fo);
    L.launches++;
    LCU(cudaGetLastError());
    short ix[4][576];
    GrInfoOut gi[4];
    unsigned char sf[4][40];
    FrameOut fo;
    LCU(cudaMemcpy(ix, L.d_ix, sizeof(ix), cudaMemcpyDeviceToHost));
    LCU(cudaMemcpy(gi, L.d_gi, sizeof(gi), cudaMemcpyDeviceToHost));
    LCU(cudaMemcpy(sf, L.d_sf, sizeof(sf), cudaMemcpyDeviceToHost));
    LCU(cudaMemcpy(&fo, L.d_fo, sizeof(fo), cudaMemcpyDeviceToHost));
    l3_side->resvDrain = fo.resv_drain;                                       // reservoir.c:223
    for (int ch = 0; ch < stereo; ch++)
        for (int b = 0; b < 4; b++) l3_side->scfsi[ch][b] = fo.scfsi[ch][b];  // loop.c:700-716
    for (int gr = 0; gr < 2; gr++)
        for (int ch = 0; ch < stereo; ch++) {
            const int g = gr * stereo + ch;
            const GrInfoOut &o = gi[g];
            gr_info *c = &l3_side->gr[gr].ch[ch].tt;
            for (int i = 0; i < 576; i++) l3_enc[gr][ch][i] = ix[g][i] < 0 ? -ix[g][i] : ix[g][i];   // magnitudes; sign: l3bitstream.c:115-125
            c->part2_3_length = o.part2_3_length; c->big_values = o.big_values; c->count1 = o.count1;
            c->global_gain = o.global_gain; c->scalefac_compress = o.scalefac_compress;
            c->table_select[0] = o.table_select[0]; c->table_select[1] = o.table_select[1]; c->table_select[2] = o.table_select[2];
            c->subblock_gain[0] = c->subblock_gain[1] = c->subblock_gain[2] = 0;
            c->region0_count = o.region0_count; c->region1_count = o.region1_count;
            c->preflag = o.preflag; c->scalefac_scale = o.scalefac_scale; c->count1table_select = o.count1table_select;
            c->part2_length = o.part2_length;
            const bool is_short = o.block_type == 2;                          // gr_deco, loop.c:2063-2081
            c->sfb_lmax = is_short ? 0 : 21; c->sfb_smax = is_short ? 0 : 12;
            c->address1 = o.address1; c->address2 = o.address2; c->address3 = o.address3;
            c->quantizerStepSize = (double)(o.global_gain - 210);
            c->sfb_partition_table = g_legacy_partition_table;
            c->slen[0] = c->slen[1] = c->slen[2] = c->slen[3] = 0;
            for (int b = 0; b < 22; b++) scalefac->l[gr][ch][b] = 0;
            for (int b = 0; b < 13; b++)
                for (int wdw = 0; wdw < 3; wdw++) scalefac->s[gr][ch][b][wdw] = 0;
            if (!is_short) for (int b = 0; b < 21; b++) scalefac->l[gr][ch][b] = sf[g][b];
            else for (int b = 0; b < 12; b++)
                for (int wdw = 0; wdw < 3; wdw++) scalefac->s[gr][ch][b][wdw] = sf[g][3 * b + wdw];
        }
}


// ---- the inner-loop functions the north star names, single-call legacy form (loop-pvt.h:27-117) ----------------
// quantize() and count_bits() have no sampling-frequency argument: like the reference (loop.c:242-250 sets the file-level
// scalefac_band_long/short in iteration_loop) they use the tables of the last iteration_loop() call, 44.1 kHz before any.
static bool legacy_rate_tables(LegacyState &L)
{
    if (L.loop_sr >= 0 && L.d_rate_tab) return true;
    RateTables *R = new RateTables;
    build_rate_tables(1, R);
    cudaError_t e = L.d_rate_tab ? cudaSuccess : cudaMalloc(&L.d_rate_tab, sizeof(RateTables));
    if (e == cudaSuccess) e = cudaMemcpy(L.d_rate_tab, R, sizeof(RateTables), cudaMemcpyHostToDevice);
    delete R;
    if (e != cudaSuccess) { legacy_fail(MP3GPU_ECUDA, "rate tables", e); return false; }
    L.loop_sr = 1;
    return true;
}

struct LegacyProbe { double *d_x = nullptr; int *d_q = nullptr, *d_bt = nullptr, *d_bits = nullptr; };
static LegacyProbe g_probe;

static bool legacy_probe_buffers()
{
    if (g_probe.d_x) return true;
    if (cudaMalloc(&g_probe.d_x, 576 * sizeof(double)) != cudaSuccess || cudaMalloc(&g_probe.d_q, sizeof(int)) != cudaSuccess ||
        cudaMalloc(&g_probe.d_bt, sizeof(int)) != cudaSuccess || cudaMalloc(&g_probe.d_bits, sizeof(int)) != cudaSuccess)
        legacy_fatal("out of device memory");
    cudaFuncSetAttribute(k_quantize_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RL_SMEM_BYTES);
    return true;
}

static void legacy_count_result(const GrInfoOut &o, gr_info *c)
{
    c->big_values = o.big_values; c->count1 = o.count1; c->count1table_select = o.count1table_select;
    c->region0_count = o.region0_count; c->region1_count = o.region1_count;
    c->table_select[0] = o.table_select[0]; c->table_select[1] = o.table_select[1]; c->table_select[2] = o.table_select[2];
    c->address1 = o.address1; c->address2 = o.address2; c->address3 = o.address3;
}

// loop.c:1360-1428: ix[i] = pow_nint(fabs(xr[i]) / 2^(quantizerStepSize/4)); subblock_gain 0, no mixed blocks
extern "C" void quantize(double xr[576], int ix[576], gr_info *cod_info)
{
    if (!legacy_init()) return;
    LegacyState &L = g_legacy;
    if (!legacy_rate_tables(L) || !legacy_probe_buffers()) return;
    double ax[576];
    for (int i = 0; i < 576; i++) ax[i] = fabs(xr[i]);
    const int q = (int)cod_info->quantizerStepSize;
    const int bt = cod_info->window_switching_flag ? (int)cod_info->block_type : 0;
    LCU(cudaMemcpy(g_probe.d_x, ax, sizeof(ax), cudaMemcpyHostToDevice));
    LCU(cudaMemcpy(g_probe.d_q, &q, sizeof(int), cudaMemcpyHostToDevice));
    LCU(cudaMemcpy(g_probe.d_bt, &bt, sizeof(int), cudaMemcpyHostToDevice));
    k_quantize_count<<<1, RL_WARPS * 32, RL_SMEM_BYTES>>>(L.d_rate_tab, g_probe.d_x, g_probe.d_q, g_probe.d_bt, 1, L.d_ix, L.d_gi, g_probe.d_bits, 0);
    L.launches++;
    LCU(cudaGetLastError());
    short out[576];
    LCU(cudaMemcpy(out, L.d_ix, sizeof(out), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 576; i++) ix[i] = out[i];
}

// loop.c:2099-2113: bits of the quantised granule; sets big_values, count1, count1table_select, region0/1_count,
// table_select[3] and address1..3 in cod_info exactly like calc_runlen / count1_bitcount / subdivide / bigv_tab_select do
extern "C" int count_bits(int *ix, gr_info *cod_info)
{
    if (!legacy_init()) return 0;
    LegacyState &L = g_legacy;
    if (!legacy_rate_tables(L) || !legacy_probe_buffers()) return 0;
    short in[576];
    for (int i = 0; i < 576; i++) in[i] = (short)(ix[i] < 0 ? -ix[i] : ix[i]);
    const int bt = cod_info->window_switching_flag ? (int)cod_info->block_type : 0;
    GrInfoOut g;
    memset(&g, 0, sizeof(g));
    g.address1 = (int)cod_info->address1; g.address2 = (int)cod_info->address2; g.address3 = (int)cod_info->address3;
    int bits = 0;
    cudaError_t e = cudaMemcpy(L.d_ix, in, sizeof(in), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(L.d_gi, &g, sizeof(g), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(g_probe.d_bt, &bt, sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        k_quantize_count<<<1, RL_WARPS * 32, RL_SMEM_BYTES>>>(L.d_rate_tab, nullptr, nullptr, g_probe.d_bt, 1, L.d_ix, L.d_gi, g_probe.d_bits, 1);
        L.launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(&g, L.d_gi, sizeof(g), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(&bits, g_probe.d_bits, sizeof(int), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { legacy_fail(MP3GPU_ECUDA, "count_bits", e); return 0; }
    legacy_count_result(g, cod_info);
    return bits;
}


// ---- the rest of the reference's inner-loop boundary (loop-pvt.h:46-117, loop.c:51-53), single-call legacy form -------------
// inner_loop, bin_search_StepSize, calc_runlen, count1_bitcount, subdivide, bigv_tab_select, new_choose_table, bigv_bitcount.
// Each reads and writes the fields of the caller's gr_info exactly as the reference's function does (no more, no fewer),
// so a host program can replace any subset of them.  One warp runs the same building blocks as the batched rate loop
// (quantize_all / count_all / count_regions, rate_loop_core.h); there is no host-side arithmetic.
namespace mp3gpu {

enum { LL_RUNLEN = 1, LL_COUNT1, LL_SUBDIVIDE, LL_TABSEL, LL_BIGV, LL_CHOOSE, LL_INNER, LL_BINSEARCH };

struct LegacyLoopArgs {
    int mode, bt, wsf;          // block type as the rate loop sees it: wsf ? block_type : 0
    int a, b;                   // max_bits | desired_rate | begin ;  start step | end
    GrInfoOut g;                // in / out: big_values, count1, count1table_select, region0/1_count, table_select, address1..3
    int result, q;              // return value, final quantizerStepSize
};

// bits of the pairs of ix[start, end) in Huffman table t, count_bit() / HuffmanCode() in count mode (loop.c:172-225)
__device__ __forceinline__ int ll_count_bit(const RateTables &T, const short *ix, int start, int end, int t, int lane)
{
    if (t == 0) return 0;
    const int ylen = T.hxlen[t], lin = T.hlinbits[t];
    const unsigned char *hl = T.hlen + T.hoff[t];
    int sum = 0;
    for (int i = start + 2 * lane; i < end; i += 64) {
        int x = ix[i], y = (i + 1 < 576) ? ix[i + 1] : 0;
        if (t > 15) {
            if (x > 14) { x = 15; sum += lin; }
            if (y > 14) { y = 15; sum += lin; }
        }
        sum += hl[x * ylen + y] + (x != 0) + (y != 0);
    }
    return __reduce_add_sync(0xffffffffu, sum);
}

__global__ void __launch_bounds__(32)
k_legacy_loop(const RateTables *__restrict__ gT, LegacyLoopArgs *A, const double *xr_abs, short *ix)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const RateHot &H = load_rate_hot(gT, smem_raw);
    RateWarpSmem &M = *reinterpret_cast<RateWarpSmem *>(smem_raw + RL_HOT_BYTES);
    const RateTables &T = *gT;
    WarpCtx w;
    const int lane = w.lane;
    LegacyLoopArgs a = *A;
    const bool is_short = a.wsf && a.bt == 2, wsf = a.wsf != 0;
    CountResult C;
    memset(&C, 0, sizeof(C));
    C.big_values = a.g.big_values; C.count1 = a.g.count1; C.count1table_select = a.g.count1table_select;
    C.region0_count = a.g.region0_count; C.region1_count = a.g.region1_count;
    C.table_select[0] = a.g.table_select[0]; C.table_select[1] = a.g.table_select[1]; C.table_select[2] = a.g.table_select[2];
    C.address1 = a.g.address1; C.address2 = a.g.address2; C.address3 = a.g.address3;
    int result = 0, q = a.b;
    // quantised values -> slots (the layouts of rate_loop_core.h); LL_CHOOSE: slot s = elements (begin + 2 s, begin + 2 s + 1)
    PerThread<int> nzmax, bigmax;
    if (a.mode == LL_RUNLEN || a.mode == LL_COUNT1 || a.mode == LL_TABSEL || a.mode == LL_CHOOSE) {
        int nz = -1, bg = -1;
        for (int k = 0; k < 9; k++) {
            const int s = lane + 32 * k;
            int e0, e1;
            if (a.mode == LL_CHOOSE) { e0 = a.a + 2 * s; e1 = e0 + 1; if (e0 >= a.b) e0 = e1 = 576; }
            else { e0 = slot_e0(is_short, s); e1 = is_short ? e0 + 3 : e0 + 1; }
            const int x = e0 < 576 ? ix[e0] : 0, y = e1 < 576 ? ix[e1] : 0;
            U2 v; v.x = (unsigned short)x; v.y = (unsigned short)y;
            M.ix[s] = v;
            if ((x | y) != 0) nz = s;
            if (x > 1 || y > 1) bg = s;
        }
        nzmax.v = nz; bigmax.v = bg;
        __syncwarp();
    }
    switch (a.mode) {
    case LL_RUNLEN:                                              // calc_runlen, loop.c:1488-1519
        if (is_short) { C.count1 = 0; C.big_values = 288; }
        else {
            const int n = w.reduce_max(nzmax) + 1, B = w.reduce_max(bigmax);
            C.count1 = (n - 1 - B) >> 1;
            C.big_values = n - 2 * C.count1;
        }
        break;
    case LL_COUNT1: {                                            // count1_bitcount, loop.c:1531-1590: the caller's big_values / count1
        int acc = 0;
        for (int t = lane; t < C.count1; t += 32) {
            const int s0 = C.big_values + 2 * t;
            if (s0 + 1 < 288) {
                const U2 u = M.ix[s0], v = M.ix[s0 + 1];
                acc += (int)H.c1lut[(u.x & 1) | ((u.y & 1) << 1) | ((v.x & 1) << 2) | ((v.y & 1) << 3)];
            }
        }
        const int both = __reduce_add_sync(0xffffffffu, acc);
        const int sum0 = both & 0xffff, sum1 = both >> 16;
        if (sum0 < sum1) { result = sum0; C.count1table_select = 0; } else { result = sum1; C.count1table_select = 1; }
        break;
    }
    case LL_SUBDIVIDE:                                           // subdivide, loop.c:1638-1704
        if (C.big_values == 0) { C.region0_count = 0; C.region1_count = 0; }
        else if (!wsf) {
            const int bv = C.big_values > 288 ? 288 : C.big_values;
            C.region0_count = H.subdiv[bv][0]; C.region1_count = H.subdiv[bv][1];
            C.address1 = H.sfb_l[C.region0_count + 1];
            C.address2 = H.sfb_l[C.region0_count + C.region1_count + 2];
            C.address3 = 2 * C.big_values;
        } else if (a.bt == 2) { C.region0_count = 8; C.region1_count = 36; C.address1 = 36; C.address2 = 2 * C.big_values; C.address3 = 0; }
        else { C.region0_count = 7; C.region1_count = 13; C.address1 = H.sfb_l[8]; C.address2 = 2 * C.big_values; C.address3 = 0; }
        break;
    case LL_TABSEL: {                                            // bigv_tab_select, loop.c:1717-1775
        CountResult D = C;
        D.table_select[0] = D.table_select[1] = D.table_select[2] = 0;
        if (is_short) { D.address1 = 36; D.address2 = 576; }     // the short branch partitions by line < 12, not by the addresses
        count_regions(w, H, M, is_short, is_short ? 576 : 2 * C.big_values, 0, D);
        C.table_select[0] = D.table_select[0]; C.table_select[1] = D.table_select[1]; C.table_select[2] = D.table_select[2];
        break;
    }
    case LL_CHOOSE: {                                            // new_choose_table(ix, begin, end), loop.c:1793-1900
        CountResult D;
        memset(&D, 0, sizeof(D));
        const int n = a.b > a.a ? (a.b - a.a + 1) / 2 : 0;
        D.address1 = 2 * (n > 288 ? 288 : n); D.address2 = 0;
        count_regions(w, H, M, false, 0, 0, D);
        result = D.table_select[0];
        break;
    }
    case LL_BIGV:                                                // bigv_bitcount, loop.c:1954-2016: the caller's tables and addresses
        if (is_short) {
            // lines < 12 of every window (elements < 36) with table_select[0], the rest with table_select[1]; a pair is
            // (line, line + 1) of one window: elements (3 l + w, 3 l + 3 + w)
            int sum = 0;
            for (int p = lane; p < 288; p += 32) {
                const int e0 = slot_e0(true, p), t = C.table_select[e0 < 36 ? 0 : 1];
                if (t == 0) continue;
                int x = ix[e0], y = ix[e0 + 3];
                const int ylen = T.hxlen[t], lin = T.hlinbits[t];
                if (t > 15) { if (x > 14) { x = 15; sum += lin; } if (y > 14) { y = 15; sum += lin; } }
                sum += T.hlen[T.hoff[t] + x * ylen + y] + (x != 0) + (y != 0);
            }
            result = __reduce_add_sync(0xffffffffu, sum);
        } else {
            result = ll_count_bit(T, ix, 0, C.address1, C.table_select[0], lane) +
                     ll_count_bit(T, ix, C.address1, C.address2, C.table_select[1], lane) +
                     ll_count_bit(T, ix, C.address2, C.address3, C.table_select[2], lane);
        }
        break;
    case LL_INNER:
    case LL_BINSEARCH: {
        for (int k = 0; k < 9; k++) {
            const int s = lane + 32 * k;
            const int e0 = slot_e0(is_short, s), e1 = is_short ? e0 + 3 : e0 + 1;
            D2 x; x.x = xr_abs[e0]; x.y = xr_abs[e1];
            M.xs[s] = x;
        }
        PerThread<float> rowmax;
        refresh_pow34(w, M, rowmax);
        C.kz = 9;
        int bits;
        if (a.mode == LL_INNER) {                                // inner_loop, loop.c:569-606 (max_bits = a.a, start step = a.b)
            for (;;) {
                bits = probe(w, H, T, M, is_short, wsf, q, rowmax, C);
                if (!(bits > a.a && q < 1024)) break;
                q += 1;
            }
        } else {                                                 // bin_search_StepSize, loop.c:2119-2140 (desired_rate = a.a, start = a.b)
            int top = a.b, bot = 200, next = a.b, last;
            do {
                last = next;
                next = (top + bot) / 2;                          // aint((top + bot) / 2.0)
                bits = probe(w, H, T, M, is_short, wsf, next, rowmax, C);
                if (bits > a.a) top = next; else bot = next;
            } while (bits != a.a && (last - next > 1 || next - last > 1));
            q = next;
        }
        result = bits;
        __syncwarp();
        for (int k = 0; k < 9; k++) {
            const int s = lane + 32 * k;
            const int e0 = slot_e0(is_short, s), e1 = is_short ? e0 + 3 : e0 + 1;
            const U2 v = M.ix[s];
            ix[e0] = (short)v.x; ix[e1] = (short)v.y;
        }
        break;
    }
    }
    if (lane == 0) {
        a.g.big_values = C.big_values; a.g.count1 = C.count1; a.g.count1table_select = C.count1table_select;
        a.g.region0_count = C.region0_count; a.g.region1_count = C.region1_count;
        a.g.table_select[0] = C.table_select[0]; a.g.table_select[1] = C.table_select[1]; a.g.table_select[2] = C.table_select[2];
        a.g.address1 = C.address1; a.g.address2 = C.address2; a.g.address3 = C.address3;
        a.result = result; a.q = q;
        *A = a;
    }
}

}  // namespace mp3gpu

static LegacyLoopArgs *g_ll_args = nullptr;

// run one mode of k_legacy_loop on the caller's gr_info; xr (576 doubles, any sign) and ix (576 ints, any sign) may be NULL
static bool legacy_loop_call(int mode, gr_info *cod, int a, int b, const double *xr, int *ix, bool ix_in, bool ix_out, LegacyLoopArgs *res)
{
    if (!legacy_init()) return false;
    LegacyState &L = g_legacy;
    if (!legacy_rate_tables(L) || !legacy_probe_buffers()) return false;
    if (!g_ll_args) {
        if (cudaMalloc(&g_ll_args, sizeof(LegacyLoopArgs)) != cudaSuccess) legacy_fatal("out of device memory");
        cudaFuncSetAttribute(k_legacy_loop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(RL_HOT_BYTES + sizeof(RateWarpSmem)));
    }
    LegacyLoopArgs h;
    memset(&h, 0, sizeof(h));
    h.mode = mode; h.a = a; h.b = b;
    if (cod) {
        h.wsf = cod->window_switching_flag != 0; h.bt = h.wsf ? (int)cod->block_type : 0;
        h.g.big_values = (int)cod->big_values; h.g.count1 = (int)cod->count1; h.g.count1table_select = (int)cod->count1table_select;
        h.g.region0_count = (int)cod->region0_count; h.g.region1_count = (int)cod->region1_count;
        h.g.table_select[0] = (int)cod->table_select[0]; h.g.table_select[1] = (int)cod->table_select[1]; h.g.table_select[2] = (int)cod->table_select[2];
        h.g.address1 = (int)cod->address1; h.g.address2 = (int)cod->address2; h.g.address3 = (int)cod->address3;
    }
    cudaError_t e = cudaMemcpy(g_ll_args, &h, sizeof(h), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && xr) {
        double ax[576];
        for (int i = 0; i < 576; i++) ax[i] = fabs(xr[i]);
        e = cudaMemcpy(g_probe.d_x, ax, sizeof(ax), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess && ix && ix_in) {
        short in[576];
        for (int i = 0; i < 576; i++) { int v = ix[i] < 0 ? -ix[i] : ix[i]; in[i] = (short)(v > 32767 ? 32767 : v); }
        e = cudaMemcpy(L.d_ix, in, sizeof(in), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) {
        k_legacy_loop<<<1, 32, RL_HOT_BYTES + sizeof(RateWarpSmem)>>>(L.d_rate_tab, g_ll_args, g_probe.d_x, L.d_ix);
        L.launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(res, g_ll_args, sizeof(*res), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && ix && ix_out) {
        short out[576];
        e = cudaMemcpy(out, L.d_ix, sizeof(out), cudaMemcpyDeviceToHost);
        for (int i = 0; i < 576; i++) ix[i] = out[i];
    }
    if (e != cudaSuccess) { legacy_fail(MP3GPU_ECUDA, "loop function", e); return false; }
    return true;
}

// loop.c:1488-1519: sets count1 and big_values
extern "C" void calc_runlen(int ix[576], gr_info *cod_info)
{
    LegacyLoopArgs r;
    if (!legacy_loop_call(LL_RUNLEN, cod_info, 0, 0, nullptr, ix, true, false, &r)) return;
    cod_info->count1 = r.g.count1; cod_info->big_values = r.g.big_values;
}

// loop.c:1531-1590: bits of the count1 quads for the caller's big_values / count1; sets count1table_select
extern "C" int count1_bitcount(int ix[576], gr_info *cod_info)
{
    LegacyLoopArgs r;
    if (!legacy_loop_call(LL_COUNT1, cod_info, 0, 0, nullptr, ix, true, false, &r)) return 0;
    cod_info->count1table_select = r.g.count1table_select;
    return r.result;
}

// loop.c:1638-1704: region0/1_count and (unless big_values == 0) address1..3 from big_values and the block type
extern "C" void subdivide(gr_info *cod_info)
{
    LegacyLoopArgs r;
    if (!legacy_loop_call(LL_SUBDIVIDE, cod_info, 0, 0, nullptr, nullptr, false, false, &r)) return;
    cod_info->region0_count = r.g.region0_count; cod_info->region1_count = r.g.region1_count;
    cod_info->address1 = r.g.address1; cod_info->address2 = r.g.address2; cod_info->address3 = r.g.address3;
}

// loop.c:1717-1775: table_select[3] for the caller's addresses / big_values
extern "C" void bigv_tab_select(int ix[576], gr_info *cod_info)
{
    LegacyLoopArgs r;
    if (!legacy_loop_call(LL_TABSEL, cod_info, 0, 0, nullptr, ix, true, false, &r)) return;
    cod_info->table_select[0] = r.g.table_select[0]; cod_info->table_select[1] = r.g.table_select[1]; cod_info->table_select[2] = r.g.table_select[2];
}

// loop.c:1793-1900: the cheapest table for ix[begin, end)
extern "C" int new_choose_table(int ix[576], unsigned int begin, unsigned int end)
{
    LegacyLoopArgs r;
    if (begin > 576) begin = 576;
    if (end > 576) end = 576;
    if (!legacy_loop_call(LL_CHOOSE, nullptr, (int)begin, (int)end, nullptr, ix, true, false, &r)) return 0;
    return r.result;
}

// loop.c:1954-2016: bits of the big-value regions with the caller's tables and addresses
extern "C" int bigv_bitcount(int ix[576], gr_info *gi)
{
    LegacyLoopArgs r;
    if (!legacy_loop_call(LL_BIGV, gi, 0, 0, nullptr, ix, true, false, &r)) return 0;
    return r.result;
}

static void legacy_loop_result(const LegacyLoopArgs &r, gr_info *c)
{
    legacy_count_result(r.g, c);
    c->quantizerStepSize = (double)r.q;
}

// loop.c:569-606: from quantizerStepSize upwards until the granule fits max_bits; ix, the count fields and the step are updated
extern "C" int inner_loop(double xr[2][2][576], int l3_enc[2][2][576], int max_bits, gr_info *cod_info, int gr, int ch)
{
    LegacyLoopArgs r;
    if (!legacy_loop_call(LL_INNER, cod_info, max_bits, (int)cod_info->quantizerStepSize, xr[gr][ch], l3_enc[gr][ch], false, true, &r)) return 0;
    legacy_loop_result(r, cod_info);
    return r.result;
}

// loop.c:2119-2140: binary search of the step between `start` and 200; returns the last step probed
extern "C" int bin_search_StepSize(int desired_rate, double start, int *ix, double xrs[576], gr_info *cod_info)
{
    LegacyLoopArgs r;
    if (!legacy_loop_call(LL_BINSEARCH, cod_info, desired_rate, (int)start, xrs, ix, false, true, &r)) return 0;
    legacy_loop_result(r, cod_info);
    return r.q;
}
