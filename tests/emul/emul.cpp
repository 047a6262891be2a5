// TEST HARNESS (host emulation of the SIMT kernels, see csrc/simt.h).  Compiled by tests/emul/Makefile
// with -DMP3GPU_HOST_EMUL; loaded by tests/test_emul_*.py.  Not part of libmp3gpu.so.
#include <stdlib.h>
#include <string.h>

#include <vector>

#define MP3GPU_RL_STATS 1
#include <ucontext.h>

#include <functional>

#include "../../mp3-enc-bsd_b200/csrc/fft_regs.h"
#include "../../mp3-enc-bsd_b200/csrc/front_core.h"
#include "../../mp3-enc-bsd_b200/csrc/psy_core.h"
#include "../../mp3-enc-bsd_b200/csrc/rate_loop_core.h"
#include "../../mp3-enc-bsd_b200/csrc/tables.h"

namespace simt { thread_local int g_tid = 0; }
namespace mp3gpu { long g_rl_stats[16]; }
extern "C" long *emul_rl_stats() { return mp3gpu::g_rl_stats; }
using namespace mp3gpu;


// ---- the 32 lanes of a warp as fibers: per-lane device code (fft_regs.h) runs unchanged; lanes::shfl / lanes::sync are
// rendezvous points of the lanes named in the mask (a divergent group waits for its own members only) ----------------
namespace {
struct WarpFibers {
    enum { RUN = 0, WAIT = 1, DONE = 2 };
    ucontext_t main_ctx, ctx[32];
    std::vector<char> stack[32];
    int state[32], src[32], cur = -1;
    unsigned want[32];
    float val[32], res[32];
    std::function<void(int)> body;
    bool deadlock = false;
};
thread_local WarpFibers *g_warp = nullptr;

void fiber_entry(int lane)
{
    WarpFibers *W = g_warp;
    W->body(lane);
    W->state[lane] = WarpFibers::DONE;
    swapcontext(&W->ctx[lane], &W->main_ctx);
}

bool run_warp(const std::function<void(int)> &body)
{
    WarpFibers W;
    W.body = body;
    g_warp = &W;
    for (int l = 0; l < 32; l++) {
        W.stack[l].resize(256 * 1024);
        getcontext(&W.ctx[l]);
        W.ctx[l].uc_stack.ss_sp = W.stack[l].data(); W.ctx[l].uc_stack.ss_size = W.stack[l].size(); W.ctx[l].uc_link = &W.main_ctx;
        makecontext(&W.ctx[l], (void (*)())fiber_entry, 1, l);
        W.state[l] = WarpFibers::RUN;
    }
    for (;;) {
        bool any = false, all_done = true;
        for (int l = 0; l < 32; l++)
            if (W.state[l] == WarpFibers::RUN) { any = true; W.cur = l; swapcontext(&W.main_ctx, &W.ctx[l]); }
        // release every group whose members have all arrived with the same mask
        for (int l = 0; l < 32; l++) {
            if (W.state[l] != WarpFibers::WAIT) continue;
            const unsigned m = W.want[l];
            bool ready = true;
            for (int k = 0; k < 32; k++) if ((m >> k & 1) && !(W.state[k] == WarpFibers::WAIT && W.want[k] == m)) ready = false;
            if (!ready) continue;
            for (int k = 0; k < 32; k++) if (m >> k & 1) W.res[k] = W.val[W.src[k]];
            for (int k = 0; k < 32; k++) if (m >> k & 1) W.state[k] = WarpFibers::RUN;
            any = true;
        }
        for (int l = 0; l < 32; l++) if (W.state[l] != WarpFibers::DONE) all_done = false;
        if (all_done) break;
        if (!any) { W.deadlock = true; break; }
    }
    g_warp = nullptr;
    return !W.deadlock;
}
}  // namespace

namespace mp3gpu {
FftRegsConst c_fftr;
namespace lanes {
float shfl(unsigned mask, float v, int src)
{
    WarpFibers *W = g_warp;
    const int l = W->cur;
    W->val[l] = v; W->src[l] = src; W->want[l] = mask; W->state[l] = WarpFibers::WAIT;
    swapcontext(&W->ctx[l], &W->main_ctx);
    return W->res[l];
}
void sync() { (void)shfl(0xffffffffu, 0.f, g_warp->cur); }
}  // namespace lanes
}  // namespace mp3gpu

extern "C" {

// FFT program vs the oracle FFT: x[n] in, transformed in place into LOGICAL order
int emul_fft(float *x, int n, int *n_ops, int *n_levels)
{
    std::vector<FftTwiddle> tw; std::vector<int> base;
    build_fft_twiddles(&tw, &base);
    FftProgram P;
    build_fft_program(n == 1024 ? 10 : 8, base, &P);
    std::vector<float> buf(x, x + n);
    run_fft_program_host(P, tw, buf.data());
    for (int i = 0; i < n; i++) x[i] = P.out_neg[i] ? -buf[P.out_slot[i]] : buf[P.out_slot[i]];
    *n_ops = (int)P.ops.size(); *n_levels = ((int)P.level_start.size() - 1) / FFT_CLASSES;
    // level sanity: ops of one level must touch disjoint slots
    for (size_t l = 0; l + FFT_CLASSES < P.level_start.size(); l += FFT_CLASSES) {   // operand-class segments of one level
        std::vector<char> used(n, 0);
        for (int i = P.level_start[l]; i < P.level_start[l + FFT_CLASSES]; i++) {
            const FftOp &o = P.ops[i];
            const uint16_t s[4] = {o.a, o.b, o.c, o.d};
            for (int j = 0; j < 4; j++) if (s[j] != 0xffff) { if (used[s[j]]) return -1; used[s[j]] = 1; }
        }
    }
    // row sanity: levels are whole rows of 32; a row is merged from at most 4 conflict-free matchings, i.e. no bank is hit
    // more than 4 times by one operand position (and full rows built from one matching are conflict free)
    for (size_t r = 0; r + 32 <= P.ops.size(); r += 32) {
        int cnt[4][32];
        memset(cnt, 0, sizeof(cnt));
        for (int i = 0; i < 32; i++) {
            const FftOp &o = P.ops[r + i];
            const uint16_t s[4] = {o.a, o.b, o.c, o.d};
            for (int j = 0; j < 4; j++)
                if (o.type != FFT_NOP && s[j] != 0xffff && ++cnt[j][FFT_SKEW((unsigned)s[j]) & 31] > 4) return -2;
        }
    }
    for (size_t l = 0; l < P.level_start.size(); l++) if (P.level_start[l] % 32) return -3;
    return 0;
}

// FFT in registers (fft_regs.h) under the fiber warp: one 1024-point and three 256-point transforms, outputs in LOGICAL order
int emul_fft_regs(const float *in_long, const float *in_short, float *out_long, float *out_short)
{
    static FftRegsPlan P;
    build_fft_regs_plan(&P);
    c_fftr = P.c;
    std::vector<float> X(FFTR_X_WORDS, 0.f);
    const bool ok = run_warp([&](int lane) {
        float xl[32], xs[3][8];
        for (int r = 0; r < 32; r++) xl[r] = in_long[lane + 32 * r];
        for (int t = 0; t < 3; t++) for (int r = 0; r < 8; r++) xs[t][r] = in_short[256 * t + lane + 32 * r];
        fft_regs_run(xl, xs, P.twA.data(), X.data(), lane);
    });
    if (!ok) return -1;
    auto val = [&](uint32_t w) { const float v = X[w & 0x7fffu]; return (w & 0x8000u) ? -v : v; };
    for (int i = 0; i <= 512; i++) { out_long[i] = val(P.out_long[i]); if (i > 0 && i < 512) out_long[1024 - i] = val(P.out_long[i] >> 16); }
    for (int t = 0; t < 3; t++)
        for (int i = 0; i <= 128; i++) {
            out_short[256 * t + i] = val(P.out_short[132 * t + i]);
            if (i > 0 && i < 128) out_short[256 * t + 256 - i] = val(P.out_short[132 * t + i] >> 16);
        }
    return 0;
}

// rate loop over one stream. xr [n_frames*2*n_ch][576], psy [same] PsyOut; outputs as the kernel writes them
int emul_rate_loop_stream(int sfreq, int n_ch, int bitrate, int n_frames, const double *xr, const PsyOut *psy,
                          short *ix, GrInfoOut *gi, unsigned char *sf, FrameOut *fo, int *max_bits)
{
    int sr = sr_index(sfreq);
    if (sr < 0) return -1;
    RateTables *T = (RateTables *)calloc(1, sizeof(RateTables));
    build_rate_tables(sr, T);
    FrameGeom G;
    frame_geometry(sfreq, n_ch, bitrate, &G);
    LoopStreamState S;
    memset(&S, 0, sizeof(S));
    static PerThread<int> st_en[4], st_xm[4];
    memset(st_en, 0, sizeof(st_en)); memset(st_xm, 0, sizeof(st_xm));
    static RateWarpSmem M;
    WarpCtx w;
    rate_loop_stream(w, T->hot, *T, M, G, S, st_en, st_xm, n_frames, xr, psy, ix, gi, sf, fo, max_bits);
    free(T);
    return 0;
}

// whole pipeline for one stream, kernel by kernel, exactly as the CUDA launch sequence does it.
// pcm: planar [n_ch][hist + n_frames*1152] with hist = 1056 leading zeros (history before the stream)
int emul_encode_stream(int sfreq, int n_ch, int bitrate, int n_frames, const short *pcm, long ch_stride, int hist,
                       double *sb /*[gc][18][32]*/, double *xr /*[gc][576]*/, PsyOut *psy /*[gc]*/, short *ix,
                       GrInfoOut *gi, unsigned char *sf, FrameOut *fo, int *max_bits)
{
    int sr = sr_index(sfreq);
    if (sr < 0) return -1;
    static FrontTables F; static PsyTables PT; static RateTables RT;
    build_front_tables(&F); build_psy_tables(sr, &PT); build_rate_tables(sr, &RT);
    std::vector<FftTwiddle> tw; std::vector<int> base;
    build_fft_twiddles(&tw, &base);
    FftProgram P10, P8;
    build_fft_program(10, base, &P10); build_fft_program(8, base, &P8);
    auto mkdev = [](const FftProgram &P, std::vector<uint32_t> &outmap) {
        outmap.assign((size_t)P.n / 2 + 1, 0u);
        auto one = [&](int i) { return (uint32_t)(FFT_SKEW((unsigned)P.out_slot[i]) | (P.out_neg[i] ? 0x8000 : 0)); };
        for (int i = 0; i <= P.n / 2; i++) outmap[i] = one(i) | ((i > 0 ? one(P.n - i) : 0u) << 16);
        FftDev d; d.words = P.words.data(); d.seg_word = P.seg_word.data(); d.n_levels = ((int)P.seg_word.size() - 1) / FFT_CLASSES; d.out = outmap.data();
        return d;
    };
    std::vector<uint32_t> o10, o8;
    PsyDev D; D.T = &PT; D.tw = tw.data(); D.f1024 = mkdev(P10, o10); D.f256 = mkdev(P8, o8);
    const int n_gran = 2 * n_frames;
    WarpCtx w;
    std::vector<PsyMid> mid((size_t)n_gran * n_ch);
    static PsyFrontSmem PF;
    for (int ch = 0; ch < n_ch; ch++)
        for (int g = 0; g < n_gran; g++) {
            memset(&mid[(size_t)g * n_ch + ch], 0, sizeof(PsyMid));
            psy_front(w, D, PF, pcm + ch * ch_stride + hist + 576L * g, &mid[(size_t)g * n_ch + ch]);
        }
    static PsyScanSmem PS;
    for (int ch = 0; ch < n_ch; ch++) {
        PsyChanState st; memset(&st, 0, sizeof(st));
        PsyScanRegs R;
        psy_scan_load(w, st, R);
        for (int g = 0; g < n_gran; g++) psy_scan_step(w, PT, PS, mid[(size_t)g * n_ch + ch], R, &psy[(size_t)g * n_ch + ch]);
        psy_scan_store(w, st, R);
    }
    static FrontWarpSmem FM;
    for (int ch = 0; ch < n_ch; ch++)
        front_walk(w, F, F.window, FM, pcm + ch * ch_stride + hist, 0, n_gran, &psy[ch].block_type, (long)(sizeof(PsyOut) / sizeof(int)) * n_ch,
                   xr + 576L * ch, 576L * n_ch, sb ? sb + 576L * ch : nullptr, 576L * n_ch);
    FrameGeom G; frame_geometry(sfreq, n_ch, bitrate, &G);
    LoopStreamState S; memset(&S, 0, sizeof(S));
    PerThread<int> st_en[4], st_xm[4];
    memset(st_en, 0, sizeof(st_en)); memset(st_xm, 0, sizeof(st_xm));
    static RateWarpSmem M;
    rate_loop_stream(w, RT.hot, RT, M, G, S, st_en, st_xm, n_frames, xr, psy, ix, gi, sf, fo, max_bits);
    return 0;
}

int emul_sizeof_psyout() { return (int)sizeof(PsyOut); }
int emul_sizeof_frameout() { return (int)sizeof(FrameOut); }
}
