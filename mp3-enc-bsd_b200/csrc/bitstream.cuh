// bitstream.cuh — Layer III bitstream formatter on the device (SURVEY §8f ranks 2+3).
//
// Replaces, for the batched API, III_format_bitstream (l3bitstream.c:68-163): encodeSideInfo (:314-458),
// encodeMainData (:179-309), Huffmancodebits (:517-716), HuffmanCode in emit mode (:779-906),
// L3_huffman_coder_count1 (:728-767) and the frame assembler BF_BitstreamFrame / WriteMainDataBits
// (formatBitstream.c:53-80, 218-247).
//
// The reference assembles the stream sequentially: main data is a continuous bit stream into which the
// header + side info of the next frame is spliced whenever a frame fills up.  Because the reference never
// pads (frame_geometry) every frame has the same size FB, so the position of every bit is known in closed
// form once the rate loop has produced part2_3_length and main_data_begin:
//     main-data byte m  ->  file byte (m / cap) * FB + SIB + m % cap        (cap = FB - SIB)
//     frame k's main data starts at main-data byte k*cap - main_data_begin[k]
// so every granule-channel can be packed and placed independently: one warp per granule-channel
// (k_bits_emit), one thread per frame for header + side info (k_bits_headers).
//
// Output goes to a per-stream sliding WINDOW of (tail + max_frames) frames: main data of a frame may
// start up to 511 main-data bytes before its own header, i.e. in frames of the previous call.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bitstream_tables.h"
#include "rate_loop_core.h"

namespace mp3gpu {

struct BitsGeom {
    int n_streams, n_frames, n_ch;
    int frame_bytes, si_bytes;      // FB, SIB
    long frame0;                    // absolute index of the call's first frame (streams advance in lockstep until their own end)
    long origin;                    // absolute file byte that window byte 0 corresponds to (may be negative)
    long wstride;                   // window bytes per stream
};

#define BITS_WARPS 8
struct alignas(16) BitsWarpSmem {
    short ix[576];
    unsigned int w[136];            // bit buffer, big-endian bit order inside each word
};

__device__ __forceinline__ void bits_or(unsigned int *w, int pos, unsigned long long code, int len)
{
    // put the low `len` bits of `code` (MSB first) at bit position pos of the buffer; len <= 57
    if (len == 0) return;
    const int word = pos >> 5, sh = pos & 31;
    // 96-bit field [word, word+3): align code so that its MSB sits at bit `sh` from the top
    const unsigned long long v = code << (64 - len);       // left-justified
    const unsigned int hi = (unsigned int)(v >> 32), lo = (unsigned int)v;
    const unsigned int a = hi >> sh;
    const unsigned int b = sh ? ((hi << (32 - sh)) | (lo >> sh)) : lo;
    const unsigned int c = sh ? (lo << (32 - sh)) : 0u;
    if (a) atomicOr(&w[word], a);
    if (b) atomicOr(&w[word + 1], b);
    if (c) atomicOr(&w[word + 2], c);
}

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int *total)
{
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    *total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
}

// code + length of big-value pair (x, y) in table t (HuffmanCode, l3bitstream.c:779-906); <= 19 + 2*(13+1) = 47 bits
__device__ __forceinline__ int huff_pair(const BitTables &T, int t, int x, int y, unsigned long long *out)
{
    if (t == 0) { *out = 0; return 0; }
    const unsigned sx = x < 0, sy = y < 0;
    x = abs(x); y = abs(y);
    const int lin = T.linbits[t];
    int xe = x, ye = y;
    if (t > 15) { xe = min(x, 15); ye = min(y, 15); }
    const unsigned e = __ldg(&T.hcode[T.hoff[t] + xe * T.ylen[t] + ye]);
    unsigned long long code = e & 0xffffffu;
    int len = (int)(e >> 24);
    if (t > 15 && x > 14) { code = (code << lin) | (unsigned)(x - 15); len += lin; }
    if (x) { code = (code << 1) | sx; len++; }
    if (t > 15 && y > 14) { code = (code << lin) | (unsigned)(y - 15); len += lin; }
    if (y) { code = (code << 1) | sy; len++; }
    *out = code;
    return len;
}

// One warp per granule-channel: scalefactors + Huffman code bits + stuffing, placed at their final position.
__global__ void __launch_bounds__(BITS_WARPS * 32)
k_bits_emit(const BitTables *__restrict__ Tg, BitsGeom G, const int *__restrict__ nfr, const short *__restrict__ ix, const GrInfoOut *__restrict__ gi,
            const unsigned char *__restrict__ sf, const FrameOut *__restrict__ fo, unsigned char *win)
{
    __shared__ BitsWarpSmem Ms[BITS_WARPS];
    const BitTables &T = *Tg;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gpf = 2 * G.n_ch;                                   // granule-channels per frame
    const long gc = (long)blockIdx.x * BITS_WARPS + warp;
    if (gc >= (long)G.n_streams * G.n_frames * gpf) return;
    const int k = (int)(gc % gpf);                                // gr * n_ch + ch
    const int gr = k / G.n_ch, ch = k % G.n_ch;
    const long sfr = gc / gpf;                                    // stream * n_frames + frame
    const int f = (int)(sfr % G.n_frames);
    const long s = sfr / G.n_frames;
    if (nfr && f >= nfr[s]) return;                               // the stream ended before this frame
    BitsWarpSmem &M = Ms[warp];
    const GrInfoOut g = gi[gc];
    const int p23 = g.part2_3_length;
    if (p23 <= 0) return;
    const FrameOut F = fo[sfr];
    // ---- position in the main-data stream -------------------------------------------------------------
    const int cap = G.frame_bytes - G.si_bytes;
    long bitpos = ((G.frame0 + f) * (long)cap - F.main_data_begin) * 8;
    for (int j = 0; j < k; j++) bitpos += gi[gc - k + j].part2_3_length;
    const int pos0 = (int)(bitpos & 7);
    // ---- stage ix, clear the bit buffer ----------------------------------------------------------------
    {
        const unsigned int *src = reinterpret_cast<const unsigned int *>(ix + gc * 576);
        unsigned int *dst = reinterpret_cast<unsigned int *>(M.ix);
#pragma unroll
        for (int i = 0; i < 9; i++) dst[lane + 32 * i] = src[lane + 32 * i];
        for (int i = lane; i < 136; i += 32) M.w[i] = 0;
    }
    __syncwarp();
    int pos = pos0;
    const bool is_short = g.window_switching_flag && g.block_type == 2;
    // ---- part 2: scalefactors (l3bitstream.c:196-251, no mixed blocks) --------------------------------
    {
        const unsigned slen1 = (0x4433322211130000ull >> (4 * g.scalefac_compress)) & 15;  // slen1_tab
        const unsigned slen2 = (0x3232132132103210ull >> (4 * g.scalefac_compress)) & 15;  // slen2_tab
        const unsigned char *sfp = sf + gc * 40;
        int len_a = 0, len_b = 0;
        unsigned va = 0, vb = 0;
        if (is_short) {
            len_a = (lane < 18) ? slen1 : slen2; va = sfp[lane];                 // e = 3*sfb + w, sfb < 6 <=> e < 18
            if (lane < 4) { len_b = slen2; vb = sfp[32 + lane]; }
        } else if (lane < 21) {
            const int band = (lane < 6) ? 0 : (lane < 11) ? 1 : (lane < 16) ? 2 : 3;
            if (gr == 0 || F.scfsi[ch][band] == 0) { len_a = (lane < 11) ? slen1 : slen2; va = sfp[lane]; }
        }
        int tot_a, tot_b;
        const int off_a = warp_excl_scan(len_a, lane, &tot_a);
        const int off_b = warp_excl_scan(len_b, lane, &tot_b);
        bits_or(M.w, pos + off_a, va, len_a);
        bits_or(M.w, pos + tot_a + off_b, vb, len_b);
        pos += tot_a + tot_b;
    }
    // ---- part 3: big values + count1 (l3bitstream.c:517-690) ------------------------------------------
    {
        const int bv = g.big_values, c1 = g.count1;
        const int r1 = T.sfb_l[g.region0_count + 1], r2 = T.sfb_l[min(g.region0_count + g.region1_count + 2, 22)];
        const int c1off = T.hoff[32 + g.count1table_select];
        unsigned long long code[9];
        int len[9], tot = 0;
#pragma unroll
        for (int j = 0; j < 9; j++) {
            const int p = 9 * lane + j;
            code[j] = 0; len[j] = 0;
            if (p < bv) {
                int t, x, y;
                if (is_short) {
                    const int e0 = T.short_e0[p];
                    t = g.table_select[p < 18 ? 0 : 1];
                    x = M.ix[e0]; y = M.ix[e0 + 3];
                } else {
                    t = g.table_select[(2 * p < r1) ? 0 : (2 * p < r2) ? 1 : 2];
                    x = M.ix[2 * p]; y = M.ix[2 * p + 1];
                }
                len[j] = huff_pair(T, t, x, y, &code[j]);
            } else if (p < bv + 2 * c1 && !((p - bv) & 1)) {
                const int v = M.ix[2 * p], w = M.ix[2 * p + 1], x = M.ix[2 * p + 2], y = M.ix[2 * p + 3];
                const int q = (v != 0) + 2 * (w != 0) + 4 * (x != 0) + 8 * (y != 0);     // |values| <= 1 here
                const unsigned e = __ldg(&T.hcode[c1off + q]);
                unsigned long long c = e & 0xffffffu;
                int l = (int)(e >> 24);
                if (v) { c = (c << 1) | (unsigned)(v < 0); l++; }
                if (w) { c = (c << 1) | (unsigned)(w < 0); l++; }
                if (x) { c = (c << 1) | (unsigned)(x < 0); l++; }
                if (y) { c = (c << 1) | (unsigned)(y < 0); l++; }
                code[j] = c; len[j] = l;
            }
            tot += len[j];
        }
        int total;
        int off = pos + warp_excl_scan(tot, lane, &total);
#pragma unroll
        for (int j = 0; j < 9; j++) { bits_or(M.w, off, code[j], len[j]); off += len[j]; }
        pos += total;
    }
    __syncwarp();
    // ---- stuffing with ones up to part2_3_length (l3bitstream.c:692-708) ------------------------------
    const int end = pos0 + p23;
    for (int wd = (pos >> 5) + lane; wd <= ((end - 1) >> 5) && pos < end; wd += 32) {
        const int lo = max(pos, wd * 32) - wd * 32, hi = min(end, wd * 32 + 32) - wd * 32;   // bit range [lo, hi) in this word
        const unsigned m = (hi - lo == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << (32 - hi));
        atomicOr(&M.w[wd], m);
    }
    __syncwarp();
    // ---- place the bytes ------------------------------------------------------------------------------
    const int nbytes = (end + 7) >> 3;
    const long m0 = bitpos >> 3;
    unsigned char *wbase = win + s * G.wstride;
    for (int i = lane; i < nbytes; i += 32) {
        const unsigned char b = (unsigned char)(M.w[i >> 2] >> (24 - 8 * (i & 3)));
        const long m = m0 + i;
        const long fm = m / cap;
        const long o = fm * G.frame_bytes + G.si_bytes + (m - fm * cap) - G.origin;
        if (o < 0 || o >= G.wstride) continue;                      // cannot happen for a consistent reservoir
        const bool shared_byte = (i == 0 && pos0 != 0) || (i == nbytes - 1 && (end & 7) != 0);
        if (shared_byte) {
            if (b) {
                unsigned char *p = wbase + o;
                const uintptr_t a = reinterpret_cast<uintptr_t>(p);
                atomicOr(reinterpret_cast<unsigned int *>(a & ~(uintptr_t)3), (unsigned int)b << (8 * (a & 3)));
            }
        } else {
            wbase[o] = b;
        }
    }
}

struct SiPacker {
    unsigned char *p; int n;
    __device__ void put(unsigned v, int bits)
    {
        for (int i = bits - 1; i >= 0; i--) { if ((v >> i) & 1u) p[n >> 3] |= (unsigned char)(0x80u >> (n & 7)); n++; }
    }
};

// One thread per frame: 32 header bits + side info (encodeSideInfo, l3bitstream.c:314-458, MPEG-1), written at the
// frame's fixed position; the thread of a stream's last frame also records the next frame's back pointer
// (formatBitstream.c:76-79) for the end-of-stream length.
__global__ void k_bits_headers(const BitTables *__restrict__ Tg, BitsGeom G, const int *__restrict__ nfr, const GrInfoOut *__restrict__ gi,
                               const FrameOut *__restrict__ fo, unsigned char *win, int *next_begin)
{
    const long sfr = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (sfr >= (long)G.n_streams * G.n_frames) return;
    const int f = (int)(sfr % G.n_frames);
    const long s = sfr / G.n_frames;
    const int n_live = nfr ? nfr[s] : G.n_frames;                 // frames of this stream in this call
    if (f >= n_live) return;
    const FrameOut F = fo[sfr];
    unsigned char buf[36];
    for (int i = 0; i < 36; i++) buf[i] = 0;
    for (int i = 0; i < 4; i++) buf[i] = Tg->header[i];
    SiPacker P{buf, 32};
    P.put((unsigned)F.main_data_begin, 9);
    P.put(0, G.n_ch == 2 ? 3 : 5);                                   // private_bits
    for (int ch = 0; ch < G.n_ch; ch++) for (int b = 0; b < 4; b++) P.put(F.scfsi[ch][b], 1);
    int mainbits = F.resv_drain;
    for (int k = 0; k < 2 * G.n_ch; k++) {
        const GrInfoOut g = gi[sfr * 2 * G.n_ch + k];
        mainbits += g.part2_3_length;
        P.put((unsigned)g.part2_3_length, 12); P.put((unsigned)g.big_values, 9); P.put((unsigned)g.global_gain, 8);
        P.put((unsigned)g.scalefac_compress, 4); P.put((unsigned)g.window_switching_flag, 1);
        if (g.window_switching_flag) {
            P.put((unsigned)g.block_type, 2); P.put((unsigned)g.mixed_block_flag, 1);
            P.put((unsigned)g.table_select[0], 5); P.put((unsigned)g.table_select[1], 5);
            P.put(0, 9);                                             // subblock_gain[3], always 0 (loop.c:318-320)
        } else {
            P.put((unsigned)g.table_select[0], 5); P.put((unsigned)g.table_select[1], 5); P.put((unsigned)g.table_select[2], 5);
            P.put((unsigned)g.region0_count, 4); P.put((unsigned)g.region1_count, 3);
        }
        P.put((unsigned)g.preflag, 1); P.put((unsigned)g.scalefac_scale, 1); P.put((unsigned)g.count1table_select, 1);
    }
    unsigned char *dst = win + s * G.wstride + ((G.frame0 + f) * (long)G.frame_bytes - G.origin);
    for (int i = 0; i < G.si_bytes; i++) dst[i] = buf[i];
    if (f == n_live - 1) next_begin[s] = F.main_data_begin + (G.frame_bytes - G.si_bytes) - mainbits / 8;
}

}  // namespace mp3gpu
