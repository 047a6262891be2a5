// tables.h — constant tables of the hot path, built ON THE HOST with the same libm expressions the
// reference uses and uploaded once per context.  Nothing here is recomputed
// with device intrinsics.
#pragma once
#include <stdint.h>

#include <vector>

#include "rate_loop_core.h"
#include "bitstream_tables.h"

namespace mp3gpu {

// ---- polyphase filterbank + MDCT (encode.c:287-409, mdct.c:25-198) ------------------------------
struct FrontTables {
    double window[512];      // Table C.1
    double am[32][32];       // am[sb][j<16] = m[sb][j], am[sb][16+j] = m[sb][33+j] (j<15); [31] unused
    double win[4][36];       // MDCT windows by block type
    double cos_l[18][36];
    double cos_s[6][12];
    double ca[8], cs[8];     // alias butterflies
    // tolerance-path variants (front_fast.cuh): the 36 -> 18 / 12 -> 6 MDCT after the time-domain aliasing fold is a DCT-IV
    double dct4_l[18][18];   // cos(pi/72 (2j+1)(2m+1)) / 9   = cos_l's kernel on the folded input
    double dct4_s[6][6];     // cos(pi/24 (2j+1)(2m+1)) / 3
};

// ---- FFT as a levelised straight-line program (subs.c:185-534) ------------------------------------
// The reference's recursive split-radix real FFT is a fixed dataflow graph.  To get bit-identical
// FP32 results on a GPU we flatten the recursion ONCE on the host into elementary in-place ops,
// schedule them into dependency levels, and let a warp execute one level per __syncwarp().
// Sign changes and data reorders (rsrec step 2/5, BR_permute) are folded into the operand maps.
enum FftOpType : uint8_t {
    FFT_BFLY = 0,   // t=a+b; b=a-b; a=t
    FFT_CROSS = 1,  // t1=a+d; t2=c+b; c=c-b; b=a-d; a=t1; d=t2
    FFT_ROT = 2,    // t2=cn*(a+c); t1=spcn*a+t2; a=smcn*c+t2; c=t1
    FFT_ROT8A = 3,  // t1=SQ*(a+c); c=SQ*(c-a); a=t1        (double multiply, rounded to float)
    FFT_ROT8B = 4,  // t2=SQ*(d-b); d=-SQ*(b+d); b=t2       (operands passed as a:=b, c:=d)
    FFT_NOP = 7,    // padding: rows of 32 ops are packed so that no two ops of a row hit the same bank
};

// Shared-memory address of logical slot s: one pad word per 32 slots, so that the power-of-two strides of the
// butterflies spread over the banks.  Device ops, the window store and the output map all use skewed addresses.
#define FFT_SKEW(s) ((s) + ((s) >> 5))
#define FFT_X_WORDS (1024 + 32)
// Padding ops of the fast classes are not skipped but aimed at per-lane dummy words behind the data (operand j of
// lane L: word FFT_X_WORDS + 32 j + L), so the executor needs no "is this a NOP" branch.
#define FFT_X_ALLOC (FFT_X_WORDS + 4 * 32)
// Batched short transforms: data set b of a 256-point batch starts FFT_BATCH_BYTES * b into x[] (264 skewed data words
// + its own 128 dummy words; 3 * 392 = 1176 <= FFT_X_ALLOC)
#define FFT_BATCH_BYTES (4 * (256 + 8 + 4 * 32))

struct FftOp {
    uint16_t a, b, c, d;  // physical slots
    uint16_t tw;          // twiddle triple index (FFT_ROT)
    uint8_t type;
    uint8_t neg;          // bit0..3: operand a,b,c,d is stored negated
    uint32_t pad;
};

// Operand classes: every row of 32 ops holds one class, executed by a specialised loop.
//   0 BFLY   butterflies without sign flips                       1 word  / op: a | b << 16
//   1 CROSS  crosses                                              2 words / op: a | b << 16, c | d << 16
//   2 ROT    twiddle rotations, operand c possibly stored negated 4 words / op: a | c << 16 | negc << 31, cn, spcn, smcn
//   3 MISC   ROT8A / ROT8B / the few butterflies with a negated a 2 words / op: a | c << 16, type | neg << 8   (b travels as c)
// Operands are BYTE offsets of skewed slots into the warp's x[] (what the shared-memory load wants).
#define FFT_CLASSES 4
#define FFT_WORDS_PAD 512   // readable words behind the last segment (look-ahead of the executor: <= 4 rows of 2 words)
inline int fft_class(const FftOp &o)
{
    if (o.type == FFT_BFLY && o.neg == 0) return 0;
    if (o.type == FFT_CROSS && o.neg == 0) return 1;
    if (o.type == FFT_ROT && (o.neg & ~4) == 0) return 2;
    return 3;
}

struct FftProgram {
    int n, logm;
    std::vector<FftOp> ops;            // sorted by (level, class), packed into bank-conflict-free rows of 32 (FFT_NOP padded)
    std::vector<uint32_t> words;       // same ops in the device encoding (see the class table above)
    std::vector<int> level_start;      // size 4*n_levels+1: ops of segment (level l, class c) = [4l+c, 4l+c+1), multiples of 32
    std::vector<int> seg_word;         // same segments as offsets into words[]
    std::vector<uint16_t> out_slot;    // logical output index -> physical slot (after bit reversal)
    std::vector<uint8_t> out_neg;      // ... stored negated?
};

struct FftTwiddle { float cn, spcn, smcn, pad; };

void build_fft_program(int logm, const std::vector<int> &tw_base, FftProgram *P);
void build_fft_twiddles(std::vector<FftTwiddle> *tw, std::vector<int> *tw_base);  // tw_base[logm]
// host interpreter (used by the emulation tests and to self-check the builder at context creation)
void run_fft_program_host(const FftProgram &P, const std::vector<FftTwiddle> &tw, float *x);

// ---- FFT in registers (fft_regs.h): the same dataflow graph, mapped on a warp without an interpreter ----------------
// A transform runs in two phases.  Phase 1 holds element j of the array in register j / 32 of lane j % 32: every
// butterfly / twiddle step of a split-radix node whose strides are >= 32 is straight-line register code, identical in
// all lanes, with the twiddle triple of line n = lane + 32 k as one coalesced 16-byte load.  The arrays then go through
// shared memory (one 36-word slot per 32-element block) and phase 2 gives every lane one whole 32-element block: the
// remaining nodes (length <= 32, plus the stride-16 steps of the length-64 complex nodes) are register code with
// compile-time twiddles; the real / imaginary arrays of a complex node sit in two lanes that exchange values by shuffle.
// The 32 + 3 x 8 blocks of a granule's 1024-point and three 256-point transforms are of four kinds and are SORTED by
// kind over the lanes of two passes (kind a: 20 + 12 blocks = one full pass of 32 lanes).
#define FFTR_SLOT_WORDS 36             // 32 data words + 4 pad words: 16-byte aligned block rows, conflict-free
#define FFTR_SLOTS 56
#define FFTR_X_WORDS (FFTR_SLOTS * FFTR_SLOT_WORDS)
enum FftBlockKind : uint8_t {
    FFTR_A = 0,   // half (real or imaginary array) of a complex node of length 32
    FFTR_B = 1,   // half of the upper half (elements 32..63) of a complex node of length 64
    FFTR_C = 2,   // real node of length 32
    FFTR_D = 3,   // elements 32..63 of a real node of length 64
};
struct FftTwC { float cn, spcn, smcn; };
struct FftSmallTw { FftTwC t4[2][4], t5[2][8], t6[2][16]; };   // [set: cn / c3n][n], nodes of length 16 / 32 / 64
// phase-1 twiddles: per node length 2^L (L = 7..10) and set, m/4 entries (cn, spcn, smcn, 0) indexed by n
#define FFTR_TWA_OFF(L, set) (((1 << ((L) - 1)) - 64) + (set) * (1 << ((L) - 2)))
#define FFTR_TWA_ENTRIES FFTR_TWA_OFF(11, 0)
struct FftRegsConst {                  // -> __constant__ memory
    FftSmallTw small;
    uint16_t long_word[32];            // first word of the slot of block r of the 1024-point array
    uint16_t short_word[3][8];         // ... of block r of short transform t
    uint8_t partner[2][32];            // [pass][lane] lane that holds the other component array (kinds a, b)
    uint8_t is_xi[2][32];              // [pass][lane] this lane holds the imaginary array
};
struct FftRegsPlan {
    FftRegsConst c;
    std::vector<float> twA;            // FFTR_TWA_ENTRIES x 4
    std::vector<uint32_t> out_long;    // bin i (0..512) -> re word | neg << 15 | (im word | neg << 15) << 16  (words into the warp's X[])
    std::vector<uint32_t> out_short;   // [3][132]
    uint8_t kind_long[32], kind_short[8];
};
void build_fft_regs_plan(FftRegsPlan *P);

// ---- psychoacoustic model tables (l3psy.c:770-994 + :194-195) -------------------------------------
struct PsyTables {
    int sr_idx, n_l, n_s;
    float hann_l[1024], hann_s[256];
    int numlines_pe[64];          // quirk: short-table counts overwrite long ones (l3psy.c:796/868)
    short part_l[513 + 3];
    short part_s[129 + 3];
    short lo_l[64], hi_l[64];     // first line / one past last line of long partition b in [0,Σlines)
    short lo_s[64], hi_s[64];
    int tail_l, tail_s;           // lines >= tail map to partition 0 (zero-initialised statics)
    double minval[64], qthr_l[64], norm_l[64];
    double qthr_s[64], norm_s[64];
    double snr_s_exp[64];         // exp(SNR_s[b] * LN_TO_LOG10), host libm (l3psy.c:712)
    double s3_l[63 * 64];         // [b*64 + k]
    double s3_lT[64 * 64];        // [k*64 + b]: the layout the kernels read (lane = b, coalesced)
    short spr_lo[64], spr_hi[64]; // spreading row range actually summed (44.1 kHz sparse, else dense+skip)
    double s3_band[64 * 64];      // [i*64 + b] = s3_lT[(spr_lo[b] + i)*64 + b]: step i of every lane's own row range is one
                                  // coalesced request, and the loop is as long as the widest range (spr_wmax), not 63
    int spr_wmax, pad_spr;
    int sparse;                   // 1 for 44.1 kHz (sprdngf1/2), 0 otherwise (dense with != 1.0 test)
    short bu_l[24], bo_l[24], bu_s[12], bo_s[12];
    double w1_l[24], w2_l[24], w1_s[12], w2_s[12];
    int n_hist_part;              // partitions that contain FFT lines 0..5 (history-dependent cw)
    // partition chains of psy_front_tail: lane l runs partition l (slot 0) and partition l + 32 (slot 1); lane 31's slot 1
    // (partition 63 does not exist) runs the lines >= tail_l that fold into partition 0
    short ch_lo[2][32], ch_hi[2][32];
    int ch_wmax[2];               // longest chain of slot 0 / slot 1
};

void build_front_tables(FrontTables *F);
void build_rate_tables(int sr_idx, RateTables *R);
void build_psy_tables(int sr_idx, PsyTables *P);
void build_bit_tables(int sr_idx, int sfreq_hz, int n_ch, int bitrate_kbps, BitTables *B);

int sr_index(int sfreq_hz);  // 0:32000 1:44100 2:48000, -1 otherwise
// frame geometry, musicin.c:562-572 and 729-746 (the reference never pads: frac_SpF is computed
// after avg_slots_per_frame was truncated)
void frame_geometry(int sfreq_hz, int n_ch, int bitrate_kbps, FrameGeom *G);

}  // namespace mp3gpu
