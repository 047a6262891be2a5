/* mp3gpu_legacy.h — the reference's own single-frame entry points, implemented by libmp3gpu.so on the GPU.
 *
 * These are the five symbols the reference's frame loop calls for the hot path
 * (/root/reference/src/musicin.c:754,767,768,774,779).  libmp3gpu.so exports them with the reference's
 * names, signatures, argument meaning and caller-visible side effects, so the reference's host program
 * (musicin.c, the PCM reader, l3bitstream.c, formatBitstream.c, the bit writer) links against the library
 * unchanged: see INTEGRATION.md for the link recipe and oracle/Makefile target `_ref/encode_gpu`.
 *
 *   symbol              reference prototype            what the shim does
 *   ------------------  -----------------------------  ---------------------------------------------
 *   window_subband      encoder.h:182 / encode.c:287   k_legacy_window: 512-sample ring + analysis window on the device
 *   filter_subband      encoder.h:184 / encode.c:361   k_legacy_filter: 8-fold + 32x31 matrixing on the device
 *   mdct_sub            mdct.h:22   / mdct.c:25        k_legacy_mdct: MDCT + alias butterflies; sign fix and slot save as the reference
 *   L3psycho_anal       l3psy.h:32  / l3psy.c:53       k_psy_front + k_psy_scan for one granule of one channel
 *   iteration_loop      loop.h:48   / loop.c:232       k_rate_loop for one frame (reservoir state on the device)
 *   quantize            loop-pvt.h  / loop.c:1360      k_quantize_count (one probe: pow_nint quantiser)
 *   count_bits          loop-pvt.h  / loop.c:2099      k_quantize_count in count-only mode (run lengths, table selection, bit count)
 *
 * Every call copies its operands to the device, launches the kernels and copies the results back
 * (one stream, one frame at a time: this is the drop-in/parity path, the batched API in mp3gpu.h is the
 * throughput path).  Like the reference these functions keep hidden state and are not reentrant, and unlike
 * the batched API they also keep the reference's error convention (musicin.c:550-557, l3psy.c:174-175): an
 * unrecoverable condition — no CUDA device, a CUDA error, an unsupported layer or sampling rate — prints a
 * message and exit(1)s, because a void function called from the unmodified musicin.c has no other way to
 * refuse; there is no CPU fallback.  mp3gpu_legacy_reset() forgets the hidden state (start of a new stream).
 *
 * The struct layouts below are the reference's (l3side.h:41-106, common.h:285-310, mdct.h:20); when the
 * reference's own headers were included first they are used instead.
 */
#ifndef MP3GPU_LEGACY_H
#define MP3GPU_LEGACY_H
#ifdef __cplusplus
extern "C" {
#endif

#ifndef L3_SIDE_H   /* l3side.h not included: declare layout-compatible types */
typedef struct {
    double l[2][2][21];
    double s[2][2][12][3];
} III_psy_ratio;

typedef struct {
    unsigned part2_3_length, big_values, count1, global_gain, scalefac_compress;
    unsigned window_switching_flag, block_type, mixed_block_flag;
    unsigned table_select[3];
    int subblock_gain[3];
    unsigned region0_count, region1_count, preflag, scalefac_scale, count1table_select;
    unsigned part2_length, sfb_lmax, sfb_smax, address1, address2, address3;
    double quantizerStepSize;
    unsigned *sfb_partition_table;
    unsigned slen[4];
} gr_info;

typedef struct {
    int main_data_begin;
    unsigned private_bits;
    int resvDrain;
    unsigned scfsi[2][4];
    struct {
        struct gr_info_s { gr_info tt; } ch[2];
    } gr[2];
} III_side_info_t;

typedef struct {
    int l[2][2][22];
    int s[2][2][13][3];
} III_scalefac_t;
#endif

#ifndef COMMON_DOT_H   /* common.h not included */
typedef struct {
    int version, lay, error_protection, bitrate_index, sampling_frequency, padding, extension, mode, mode_ext,
        copyright, original, emphasis;
} layer;
typedef struct {
    layer *header;
    int actual_mode;
    void *alloc;
    int tab_num, stereo, jsbound, sblimit;
} frame_params;
#endif

#ifndef MP3GPU_LEGACY_NO_PROTOTYPES   /* define when the reference's own (K&R) prototypes are in scope */
typedef double mp3gpu_L3SBS[2][3][18][32];
void window_subband(short **buffer, double z[512], int k);
void filter_subband(double z[512], double s[32]);
void mdct_sub(mp3gpu_L3SBS *sb_sample, double (*mdct_freq)[2][576], int stereo, III_side_info_t *l3_side, int mode_gr);
void L3psycho_anal(short *buffer, short savebuf[1344], int chn, int lay, float snr32[32], double sfreq,
                   double ratio_d[21], double ratio_ds[12][3], double *pe, gr_info *cod_info);
void iteration_loop(double pe[][2], double xr_org[2][2][576], III_psy_ratio *ratio, III_side_info_t *l3_side,
                    int l3_enc[2][2][576], int mean_bits, int stereo, double xr_dec[2][2][576],
                    III_scalefac_t *scalefac, frame_params *fr_ps, int ancillary_pad, int bitsPerFrame);
/* the inner-loop pair the rate loop is built from (loop-pvt.h:27-117, loop.c:1360 and :2099), one granule per call:
 * quantize() at cod_info->quantizerStepSize, count_bits() = calc_runlen + count1_bitcount + subdivide + bigv_tab_select
 * (new_choose_table) + bigv_bitcount, updating cod_info like the reference.  Like the reference's they use the
 * scalefactor-band tables of the last iteration_loop() call (44.1 kHz before any). */
void quantize(double xr[576], int ix[576], gr_info *cod_info);
int count_bits(int *ix, gr_info *cod_info);
/* the rest of the inner-loop boundary (loop-pvt.h:46-117, loop.c:51-53), each reading and writing exactly the gr_info fields the
 * reference's function does: inner_loop (loop.c:569: steps from quantizerStepSize upwards until the granule fits max_bits),
 * bin_search_StepSize (loop.c:2119), calc_runlen (:1488), count1_bitcount (:1531), subdivide (:1638), bigv_tab_select (:1717),
 * new_choose_table (:1793), bigv_bitcount (:1954). */
int inner_loop(double xr[2][2][576], int l3_enc[2][2][576], int max_bits, gr_info *cod_info, int gr, int ch);
int bin_search_StepSize(int desired_rate, double start, int *ix, double xrs[576], gr_info *cod_info);
void calc_runlen(int ix[576], gr_info *cod_info);
int count1_bitcount(int ix[576], gr_info *cod_info);
void subdivide(gr_info *cod_info);
void bigv_tab_select(int ix[576], gr_info *cod_info);
int new_choose_table(int ix[576], unsigned int begin, unsigned int end);
int bigv_bitcount(int ix[576], gr_info *gi);
#endif

/* forget the hidden per-stream state of all five entry points */
void mp3gpu_legacy_reset(void);
/* number of kernels the legacy entry points launched so far */
long mp3gpu_legacy_kernel_launches(void);

#ifdef __cplusplus
}
#endif
#endif
