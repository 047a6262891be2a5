/* mp3gpu_encode.c — the reference's frame loop (musicin.c:585-800) as a BATCHED C host on libmp3gpu.so.
 *
 *   mp3gpu_encode [-m s|m] [-s 44.1|48|32] [-b kbps] [-d device] [-c frames_per_call] out_dir in1.wav [in2.wav ...]
 *
 * Encodes all inputs (same sampling rate / channels / bitrate, any lengths) in ONE batch and writes out_dir/<name>.mp3.
 * Flags as the reference CLI's (musicin.c:157-378: -m mode, -s sampling frequency in kHz, -b bitrate); inputs are read as
 * the reference reads them: a file with "WAVE" at bytes 8..11 has its samples at 0x2c (musicin.c:352-368), anything else
 * is raw little-endian 16-bit PCM; num_samples = payload bytes / 2, the last frame is zero-filled (encode.c:162-166).
 * What replaces what:
 *   get_audio / read_samples (encode.c:107-269)       -> the interleaved samples go to the library as they are
 *                                                        (MP3GPU_PCM_INTERLEAVED: the channel split runs on the device)
 *   L3psycho_anal, window_subband, filter_subband,
 *   mdct_sub, iteration_loop, III_format_bitstream
 *   (musicin.c:751-785)                               -> mp3gpu_encode_frames_mp3, `frames_per_call` frames of every stream
 *   III_FlushBitstream (musicin.c:809)                -> mp3gpu_flush_mp3
 *   close_bit_stream_w (common.c:968-974)             -> the one zero byte the reference appends is written here
 * The files are byte-identical to the reference CLI's (tests/test_gpu_host_c.py).  Plain C99, no CUDA headers: pageable
 * host buffers (the library stages them; pinned memory would only make the copies asynchronous).
 * Build: gcc -O2 -std=c99 -Iinclude examples/mp3gpu_encode.c -o examples/mp3gpu_encode -Lmp3-enc-bsd_b200 -lmp3gpu \
 *        -Wl,-rpath,'$ORIGIN/../mp3-enc-bsd_b200'
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mp3gpu.h"

typedef struct {
    const char *path;
    int16_t *samples;      /* interleaved, as in the file */
    long n_samples;        /* all channels */
    long n_frames;
} input_t;

static void die(const char *what, const char *detail)
{
    fprintf(stderr, "mp3gpu_encode: %s%s%s\n", what, detail ? ": " : "", detail ? detail : "");
    exit(1);
}

static void read_input(input_t *in, int n_ch)
{
    FILE *f = fopen(in->path, "rb");
    unsigned char head[0x2c];
    long size, off = 0;
    if (!f) die("cannot open", in->path);
    fseek(f, 0, SEEK_END);
    size = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (size >= 0x2c && fread(head, 1, 0x2c, f) == 0x2c && memcmp(head + 8, "WAVE", 4) == 0) off = 0x2c;   /* musicin.c:352-362 */
    fseek(f, off, SEEK_SET);
    in->n_samples = (size - off) / 2;                                                                   /* musicin.c:368 */
    in->samples = (int16_t *)malloc((size_t)(in->n_samples > 0 ? in->n_samples : 1) * 2);
    if (!in->samples || fread(in->samples, 2, (size_t)in->n_samples, f) != (size_t)in->n_samples) die("cannot read", in->path);
    fclose(f);
    in->n_frames = (in->n_samples + 1152L * n_ch - 1) / (1152L * n_ch);                                 /* frame loop until get_audio() == 0 */
}

int main(int argc, char **argv)
{
    int n_ch = 2, sfreq = 44100, kbps = 128, device = 0, F = 30, a = 1, s, S;
    const char *out_dir;
    input_t *in;
    long max_frames = 0, *frames, *len, f0, stride;
    int16_t *pcm;
    uint8_t *mp3;
    mp3gpu_config cfg;
    mp3gpu_ctx *gpu;
    int FB;

    for (; a + 1 < argc && argv[a][0] == '-'; a += 2) {
        const char *v = argv[a + 1];
        switch (argv[a][1]) {
        case 'm': n_ch = (v[0] == 'm') ? 1 : 2; break;
        case 's': sfreq = (int)(atof(v) * 1000.0 + 0.5); break;
        case 'b': kbps = atoi(v); break;
        case 'd': device = atoi(v); break;
        case 'c': F = atoi(v); break;
        default: die("unknown flag", argv[a]);
        }
    }
    if (argc - a < 2 || F < 1) die("usage: mp3gpu_encode [-m s|m] [-s kHz] [-b kbps] [-d device] [-c frames_per_call] out_dir in.wav ...", NULL);
    out_dir = argv[a++];
    S = argc - a;
    in = (input_t *)calloc((size_t)S, sizeof(*in));
    frames = (long *)calloc((size_t)S, sizeof(long));
    len = (long *)calloc((size_t)S, sizeof(long));
    for (s = 0; s < S; s++) {
        in[s].path = argv[a + s];
        read_input(&in[s], n_ch);
        frames[s] = in[s].n_frames;
        if (frames[s] > max_frames) max_frames = frames[s];
    }

    cfg.sfreq_hz = sfreq; cfg.n_ch = n_ch; cfg.bitrate_kbps = kbps; cfg.max_streams = S; cfg.max_frames = F; cfg.device = device;
    if (mp3gpu_create(&cfg, &gpu) != MP3GPU_OK) die("mp3gpu_create", mp3gpu_last_error());   /* no CUDA device: an error, never a CPU path */
    if (mp3gpu_set_pcm_layout(gpu, MP3GPU_PCM_INTERLEAVED) || mp3gpu_set_stream_frames(gpu, S, frames, NULL) ||
        mp3gpu_frame_bytes(gpu, &FB, NULL)) die("mp3gpu setup", mp3gpu_last_error());
    stride = max_frames * FB;
    pcm = (int16_t *)malloc((size_t)S * F * 1152 * n_ch * sizeof(int16_t));
    mp3 = (uint8_t *)calloc((size_t)S * (size_t)(stride > 0 ? stride : 1), 1);
    if (!pcm || !mp3) die("out of memory", NULL);

    for (f0 = 0; f0 < max_frames; f0 += F) {
        const long nf = (max_frames - f0 < F) ? max_frames - f0 : F, per = nf * 1152 * n_ch;
        for (s = 0; s < S; s++) {                                     /* this call's samples of every stream, zero-filled past its end */
            long have = in[s].n_samples - f0 * 1152 * n_ch;
            if (have < 0) have = 0;
            if (have > per) have = per;
            memcpy(pcm + (size_t)s * per, in[s].samples + f0 * 1152 * n_ch, (size_t)have * 2);
            memset(pcm + (size_t)s * per + have, 0, (size_t)(per - have) * 2);
        }
        if (mp3gpu_encode_frames_mp3(gpu, pcm, S, (int)nf, mp3, stride, NULL)) die("mp3gpu_encode_frames_mp3", mp3gpu_last_error());
        if (mp3gpu_sync(gpu, NULL)) die("mp3gpu_sync", mp3gpu_last_error());          /* pcm is reused by the next trip */
    }
    if (mp3gpu_flush_mp3(gpu, S, mp3, stride, len, NULL)) die("mp3gpu_flush_mp3", mp3gpu_last_error());

    for (s = 0; s < S; s++) {
        char path[4096];
        const char *base = strrchr(in[s].path, '/');
        const char *dot;
        size_t n;
        FILE *f;
        base = base ? base + 1 : in[s].path;
        dot = strrchr(base, '.');
        n = dot ? (size_t)(dot - base) : strlen(base);
        snprintf(path, sizeof(path), "%s/%.*s.mp3", out_dir, (int)n, base);
        f = fopen(path, "wb");
        if (!f) die("cannot write", path);
        fwrite(mp3 + (size_t)s * stride, 1, (size_t)len[s], f);
        fputc(0, f);                                                  /* close_bit_stream_w(), common.c:968-974 */
        fclose(f);
        printf("%s: %ld frames, %ld bytes\n", path, frames[s], len[s] + 1);
    }
    printf("%ld kernel launches\n", mp3gpu_kernel_launches(gpu));
    mp3gpu_destroy(gpu);
    return 0;
}
