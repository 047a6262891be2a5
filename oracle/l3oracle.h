/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * l3oracle: a plain-C, single-threaded CPU restatement of the Layer III hot path of
 * lieff/mp3-enc-bsd (polyphase filterbank, MDCT + alias reduction, psychoacoustic model 2,
 * rate loop incl. the bit-reservoir recurrence), with the reference's function statics turned
 * into explicit per-stream state so many streams can be checked in one process.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm may build,
 * link, import or execute anything under oracle/.  The product path (libmp3gpu.so) never does.
 *
 * Parity pinning: the reference ships NO golden vectors / KATs for this path, so
 * this restatement is pinned against the UNMODIFIED reference compiled from /root/reference/src
 * (oracle/_ref/libref.so, recipe in oracle/Makefile) — tests/test_oracle_vs_ref.py requires
 * bit-identical sb/xr/pe/ratio/ix/side-info on the SURVEY §8d synthetic configs — and against
 * the fixtures committed under tests/golden/ (generated from libref.so by tools/make_golden.py).
 */
#ifndef L3ORACLE_H
#define L3ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* side info of one granule-channel; field order == oracle/ref_harness.py GR_FIELDS */
typedef struct {
    int part2_3_length, big_values, count1, global_gain, scalefac_compress;
    int window_switching_flag, block_type, mixed_block_flag;
    int table_select[3];
    int region0_count, region1_count, preflag, scalefac_scale, count1table_select;
    int part2_length, address1, address2, address3;
} l3o_gr_info;

/* everything the reference's frame loop produces for one frame (musicin.c:751-786) */
typedef struct {
    double sb[2][2][18][32];      /* [gr][ch] raw filter_subband output (before mdct sign fix) */
    double xr[2][2][576];         /* mdct_sub output */
    double pe[2][2];
    double ratio_l[2][2][21];
    double ratio_s[2][2][12][3];
    int block_type[2][2];
    int max_bits[2][2];
    int ix[2][2][576];            /* |ix| as iteration_loop leaves it */
    l3o_gr_info gi[2][2];
    double qstep[2][2];
    int scalefac_l[2][2][22];
    int scalefac_s[2][2][13][3];
    int scfsi[2][4];
    int resv_drain;
    int resv_size;                /* reservoir size after ResvFrameEnd */
} l3o_frame;

typedef struct l3o_enc l3o_enc;

/* sfreq in Hz (32000/44100/48000), n_ch 1|2, bitrate in kbps (MPEG-1 Layer III table) */
l3o_enc *l3o_create(int sfreq, int n_ch, int bitrate_kbps);
void l3o_destroy(l3o_enc *e);
/* pcm planar [n_ch][1152]; returns 0 on success */
int l3o_encode_frame(l3o_enc *e, const short *pcm, l3o_frame *out);
/* whole stream: pcm planar [n_ch][n_samples] (zero padded to whole frames); out[n_frames] */
int l3o_encode_stream(int sfreq, int n_ch, int bitrate_kbps, const short *pcm, long n_samples,
                      l3o_frame *out, long n_frames);

/* ---- stage-level entry points (stateless, for kernel-level parity tests) ---- */
/* polyphase: n_slots*32 samples of one channel, zero history; sb[n_slots][32] */
void l3o_polyphase(const short *pcm, long n_slots, double *sb);
/* one granule MDCT: prev/cur raw subband samples [18][32] (sign fix applied inside), xr[576] */
void l3o_mdct_granule(const double *prev, const double *cur, int block_type, double *xr);
/* FP32 real FFT + energy/phase as the reference's fft() (subs.c:38-123); n = 1024 or 256;
 * x is destroyed; energy/phi get n/2+1 entries */
void l3o_fft(float *x, int n, float *energy, float *phi);
/* quantize + count_bits for one granule at step q (loop.c:1360-1428, 2099-2113):
 * xr_abs[576] magnitudes, block_type, sr index 0:32k 1:44.1k 2:48k. Fills ix[576] and gi's
 * big_values,count1,count1table_select,region0/1_count,table_select,address1-3. Returns bits. */
int l3o_quantize_count(const double *xr_abs, int q, int block_type, int sr_idx, int *ix, l3o_gr_info *gi);
/* count_bits on a given ix (bit-exact Huffman bit count + table selection) */
int l3o_count_bits(const int *ix, int block_type, int sr_idx, l3o_gr_info *gi);

/* Bitstream formatter (l3bitstream.c + formatBitstream.c restated): frames -> the byte stream the reference CLI
 * writes, including the one trailing byte close_bit_stream_w() adds.  mdb[n_frames] (optional) receives
 * main_data_begin of each frame.  Returns the byte count, -1 if cap is too small. */
long l3o_format_stream(int sfreq, int n_ch, int bitrate_kbps, const l3o_frame *frames, long n_frames, unsigned char *out, long cap,
                       int *mdb);

#ifdef __cplusplus
}
#endif
#endif
