"""TEST INFRASTRUCTURE — a minimal MPEG-1 Layer III decoder (ISO 11172-3 §2.4.3.4) in numpy, for the decoded-SNR
report and to prove that streams which are NOT byte-identical to the reference's (segment seams) are
still valid Layer III.  The reference repository contains no decoder.  Supports what this encoder emits: MPEG-1,
no CRC, plain stereo / mono, long / start / stop / short blocks (no mixed blocks), constant frame size.

Only tests/ and bench tooling may import this module.  Huffman and band tables come from liboracle.so
(the same ISO tables the encoder uses, iso_tables.h)."""
import ctypes as C

import numpy as np

import oracle

_T = {}


def _tables():
    if _T:
        return _T
    L = oracle.lib()
    L.l3o_huff_table.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.l3o_sfb_tables.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    L.l3o_analysis_window.argtypes = [C.c_void_p]
    huff = {}
    for t in range(34):
        xl, yl, lb = C.c_int(), C.c_int(), C.c_int()
        codes = (C.c_uint * 256)()
        lens = (C.c_ubyte * 256)()
        n = L.l3o_huff_table(t, C.byref(xl), C.byref(yl), C.byref(lb), codes, lens)
        dec = {}
        for i in range(n):
            if lens[i]:
                dec[(lens[i], codes[i])] = (i, 0) if t >= 32 else divmod(i, yl.value)
        huff[t] = (dec, lb.value, max([k[0] for k in dec], default=0))
    sfb = {}
    for sr in range(3):
        l, s = (C.c_short * 23)(), (C.c_short * 14)()
        L.l3o_sfb_tables(sr, l, s)
        sfb[sr] = (list(l), list(s))
    win = (C.c_double * 512)()
    L.l3o_analysis_window(win)
    i = np.arange(36)
    w = np.zeros((4, 36))
    w[0] = np.sin(np.pi / 36 * (i + 0.5))
    w[1, :18] = w[0, :18]; w[1, 18:24] = 1.0; w[1, 24:30] = np.sin(np.pi / 12 * (i[24:30] - 18 + 0.5))
    w[3, 6:12] = np.sin(np.pi / 12 * (i[6:12] - 6 + 0.5)); w[3, 12:18] = 1.0; w[3, 18:] = w[0, 18:]
    w[2, :12] = np.sin(np.pi / 12 * (i[:12] + 0.5))
    k = np.arange(18)
    imdct_l = np.cos(np.pi / 72 * np.outer(2 * i + 1 + 18, 2 * k + 1))          # [36][18]
    imdct_s = np.cos(np.pi / 24 * np.outer(2 * np.arange(12) + 1 + 6, 2 * np.arange(6) + 1))  # [12][6]
    ci = np.array([-0.6, -0.535, -0.33, -0.185, -0.095, -0.041, -0.0142, -0.0037])
    cs, ca = 1 / np.sqrt(1 + ci * ci), ci / np.sqrt(1 + ci * ci)
    synth_n = np.cos(np.outer(16 + np.arange(64), 2 * np.arange(32) + 1) * np.pi / 64)   # [64][32]
    _T.update(huff=huff, sfb=sfb, D=32.0 * np.array(win[:]), win=w, imdct_l=imdct_l, imdct_s=imdct_s, cs=cs, ca=ca, N=synth_n,
              pretab=[0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 3, 2, 0])
    return _T


class _Bits:
    def __init__(self, data):
        self.b = np.unpackbits(np.frombuffer(bytes(data), dtype=np.uint8)).tolist()
        self.p = 0

    def get(self, n):
        v = 0
        for x in self.b[self.p:self.p + n]:
            v = (v << 1) | x
        self.p += n
        return v


SLEN = ([0, 0, 0, 0, 3, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4], [0, 1, 2, 3, 0, 1, 2, 3, 1, 2, 3, 1, 2, 3, 2, 3])
RATES = [0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320]
FREQS = [44100, 48000, 32000]


def _huff_pair(bits, table):
    dec, linbits, maxlen = table
    code, n = 0, 0
    while n < maxlen:
        code = (code << 1) | bits.get(1)
        n += 1
        hit = dec.get((n, code))
        if hit is not None:
            return hit, linbits
    raise ValueError("invalid Huffman code")


def parse_frames(data):
    """split a constant-frame-size MPEG-1 Layer III stream; returns (sfreq, n_ch, frame_bytes, [frame dict])"""
    T = _tables()
    h = int.from_bytes(data[:4], "big")
    assert (h >> 20) == 0xfff and ((h >> 19) & 1) == 1 and ((h >> 17) & 3) == 1, "not an MPEG-1 Layer III header"
    assert (h >> 16) & 1, "CRC not supported"
    br, sf_i, mode = RATES[(h >> 12) & 15], (h >> 10) & 3, (h >> 6) & 3
    sfreq, n_ch = FREQS[sf_i], 1 if mode == 3 else 2
    fb = int(1152 / sfreq * br * 1000 / 8)
    sib = 4 + (17 if n_ch == 1 else 32)
    frames, main = [], bytearray()
    for off in range(0, len(data), fb):
        fr = data[off:off + fb]
        if len(fr) < sib or fr[:2] != data[:2]:
            break
        b = _Bits(fr[4:sib])
        mdb = b.get(9)
        b.get(5 if n_ch == 1 else 3)
        scfsi = [[b.get(1) for _ in range(4)] for _ in range(n_ch)]
        gis = []
        for gr in range(2):
            for ch in range(n_ch):
                g = dict(part2_3_length=b.get(12), big_values=b.get(9), global_gain=b.get(8), scalefac_compress=b.get(4),
                         wsf=b.get(1))
                if g["wsf"]:
                    g["block_type"], g["mixed"] = b.get(2), b.get(1)
                    g["table_select"] = [b.get(5), b.get(5), 0]
                    g["sbg"] = [b.get(3) for _ in range(3)]
                    g["region0_count"], g["region1_count"] = (8, 36) if g["block_type"] == 2 else (7, 13)
                else:
                    g["block_type"], g["mixed"], g["sbg"] = 0, 0, [0, 0, 0]
                    g["table_select"] = [b.get(5), b.get(5), b.get(5)]
                    g["region0_count"], g["region1_count"] = b.get(4), b.get(3)
                g["preflag"], g["scalefac_scale"], g["count1table_select"] = b.get(1), b.get(1), b.get(1)
                gis.append(g)
        frames.append(dict(mdb=mdb, scfsi=scfsi, gi=gis, main_start=len(main) - mdb))
        main += fr[sib:]
    return sfreq, n_ch, fb, frames, bytes(main)


def decode_spectra(data):
    """-> (sfreq, n_ch, xr [n_frames*2][n_ch][576], ix same shape (int), ok flags per frame)"""
    T = _tables()
    sfreq, n_ch, fb, frames, main = parse_frames(data)
    sr = {32000: 0, 44100: 1, 48000: 2}[sfreq]
    sfb_l, sfb_s = T["sfb"][sr]
    bits = _Bits(main)
    nf = len(frames)
    xr = np.zeros((nf * 2, n_ch, 576))
    ixs = np.zeros((nf * 2, n_ch, 576), np.int32)
    ok = np.ones(nf, bool)
    sf_prev = [[0] * 22 for _ in range(n_ch)]
    for f, fr in enumerate(frames):
        if fr["main_start"] < 0:
            ok[f] = False       # needs main data from before the stream start (cut stream)
            continue
        pos = fr["main_start"] * 8
        for gr in range(2):
            for ch in range(n_ch):
                g = fr["gi"][gr * n_ch + ch]
                bits.p = pos
                end = pos + g["part2_3_length"]
                pos = end
                slen1, slen2 = SLEN[0][g["scalefac_compress"]], SLEN[1][g["scalefac_compress"]]
                short = g["wsf"] and g["block_type"] == 2
                sf_l, sf_s = [0] * 22, [[0] * 3 for _ in range(13)]
                if short:
                    for sfb in range(12):
                        for w in range(3):
                            sf_s[sfb][w] = bits.get(slen1 if sfb < 6 else slen2)
                else:
                    for band, (lo, hi) in enumerate(((0, 6), (6, 11), (11, 16), (16, 21))):
                        for sfb in range(lo, hi):
                            if gr == 1 and fr["scfsi"][ch][band]:
                                sf_l[sfb] = sf_prev[ch][sfb]
                            else:
                                sf_l[sfb] = bits.get(slen1 if sfb < 11 else slen2)
                    if gr == 0 or True:
                        sf_prev[ch] = list(sf_l)
                # Huffman
                ix = [0] * 578
                bv2 = min(g["big_values"] * 2, 576)
                if short:
                    r1, r2 = 36, 576
                else:
                    r1 = sfb_l[min(g["region0_count"] + 1, 22)]
                    r2 = sfb_l[min(g["region0_count"] + g["region1_count"] + 2, 22)]
                i = 0
                while i < bv2:
                    t = g["table_select"][0 if i < r1 else 1 if i < r2 else 2]
                    if t == 0:
                        i += 2
                        continue
                    (x, y), linbits = _huff_pair(bits, T["huff"][t])
                    if x == 15 and linbits:
                        x += bits.get(linbits)
                    if x and bits.get(1):
                        x = -x
                    if y == 15 and linbits:
                        y += bits.get(linbits)
                    if y and bits.get(1):
                        y = -y
                    ix[i], ix[i + 1] = x, y
                    i += 2
                tab = T["huff"][32 + g["count1table_select"]]
                while bits.p < end and i <= 572:
                    (p, _), _ = _huff_pair(bits, tab)
                    q = [(p >> k) & 1 for k in range(4)]
                    for k in range(4):
                        if q[k] and bits.get(1):
                            q[k] = -1
                    if bits.p > end:
                        break       # overshoot into stuffing: discard (2.4.3.4.6)
                    ix[i:i + 4] = q
                    i += 4
                ixa = np.array(ix[:576], np.int64)
                ixs[f * 2 + gr, ch] = ixa
                # requantisation (2.4.3.4.7.1)
                mult = 0.5 * (1 + g["scalefac_scale"])
                mag = np.sign(ixa) * np.abs(ixa) ** (4.0 / 3.0)
                if short:
                    # bitstream order is sfb / window / line; spectrum order of this encoder is line-major [192][3]
                    out = np.zeros(576)
                    k = 0
                    for sfb in range(13):
                        lo, hi = sfb_s[sfb], sfb_s[sfb + 1]
                        for w in range(3):
                            sc = 2.0 ** (0.25 * (g["global_gain"] - 210 - 8 * g["sbg"][w])) * 2.0 ** (-mult * (sf_s[sfb][w] if sfb < 12 else 0))
                            for line in range(lo, hi):
                                out[3 * line + w] = mag[k] * sc
                                k += 1
                    # ixs in spectrum order too, for comparisons with the encoder's ix
                    o2 = np.zeros(576, np.int64)
                    k = 0
                    for sfb in range(13):
                        for w in range(3):
                            for line in range(sfb_s[sfb], sfb_s[sfb + 1]):
                                o2[3 * line + w] = ixa[k]
                                k += 1
                    ixs[f * 2 + gr, ch] = o2
                    xr[f * 2 + gr, ch] = out
                else:
                    sc = np.zeros(576)
                    for sfb in range(22):
                        lo, hi = sfb_l[sfb], sfb_l[sfb + 1]
                        s = sf_l[sfb] + (g["preflag"] * T["pretab"][sfb])
                        sc[lo:hi] = 2.0 ** (0.25 * (g["global_gain"] - 210)) * 2.0 ** (-mult * s)
                    xr[f * 2 + gr, ch] = mag * sc
                fr["gi"][gr * n_ch + ch]["short"] = short
    bt = np.array([[fr["gi"][gr * n_ch + ch]["block_type"] for ch in range(n_ch)] for fr in frames for gr in range(2)])
    return sfreq, n_ch, xr, ixs, bt, ok


def synthesize(xr, bt):
    """hybrid synthesis: alias reduction, IMDCT + overlap-add, frequency inversion, polyphase synthesis.
    xr [n_gran][n_ch][576] (short blocks line-major [192][3] as decode_spectra leaves them) -> pcm float [n_ch][n_gran*576]"""
    T = _tables()
    n_gran, n_ch, _ = xr.shape
    pcm = np.zeros((n_ch, n_gran * 576))
    for ch in range(n_ch):
        prev = np.zeros((32, 18))
        V = np.zeros(1024)
        for g in range(n_gran):
            x = xr[g, ch].copy()
            t = int(bt[g, ch])
            sbs = np.zeros((32, 36))
            if t == 2:
                for sb in range(32):
                    blk = x[18 * sb:18 * sb + 18].reshape(6, 3)        # [m][window]
                    for w in range(3):
                        sbs[sb, 6 + 6 * w:18 + 6 * w] += (T["imdct_s"] @ blk[:, w]) * T["win"][2, :12]
            else:
                for sb in range(1, 32):
                    for i in range(8):
                        bu, bd = x[18 * sb - 1 - i], x[18 * sb + i]
                        x[18 * sb - 1 - i] = bu * T["cs"][i] - bd * T["ca"][i]
                        x[18 * sb + i] = bd * T["cs"][i] + bu * T["ca"][i]
                sbs = (x.reshape(32, 18) @ T["imdct_l"].T) * T["win"][t]
            out = sbs[:, :18] + prev                                       # [sb][time]
            prev = sbs[:, 18:].copy()
            out[1::2, 1::2] *= -1.0                                        # frequency inversion
            for ts in range(18):
                V = np.roll(V, 64)
                V[:64] = T["N"] @ out[:, ts]
                U = np.zeros(512)
                for i in range(8):
                    U[64 * i:64 * i + 32] = V[128 * i:128 * i + 32]
                    U[64 * i + 32:64 * i + 64] = V[128 * i + 96:128 * i + 128]
                W = U * T["D"]
                pcm[ch, g * 576 + 32 * ts:g * 576 + 32 * ts + 32] = W.reshape(16, 32).sum(axis=0)
    return pcm * 32768.0


def decode(data):
    """MP3 bytes -> (sfreq, float PCM [n_ch][n] in int16 units, per-frame ok flags)"""
    sfreq, n_ch, xr, ix, bt, ok = decode_spectra(data)
    return sfreq, synthesize(xr, bt), ok


def snr_db(ref, test):
    """10 log10(sum ref^2 / sum (ref-test)^2) over the common length"""
    n = min(ref.shape[-1], test.shape[-1])
    e = ref[..., :n] - test[..., :n]
    den = float((e * e).sum())
    return float("inf") if den == 0 else 10 * np.log10(float((ref[..., :n] ** 2).sum()) / den)


CODEC_DELAY = 1057   # samples: 481 (analysis + synthesis polyphase filterbanks) + 576 (MDCT overlap)


def snr_vs_original(orig, dec, delay=CODEC_DELAY, skip=2304):
    """SNR in dB of the decoded signal against the encoder input at the known codec delay"""
    m = min(orig.shape[1], dec.shape[1] - delay)
    return snr_db(orig[:, skip:m].astype(np.float64), dec[:, delay + skip:delay + m])


def align_and_snr(orig, dec, max_delay=2400):
    """find the codec delay by cross-correlation on channel 0, return (delay, SNR in dB of the decoded signal)"""
    a, b = orig[0].astype(np.float64), dec[0]
    n = min(len(a), len(b)) - max_delay
    seg = slice(4608, min(n, 4608 + 44100))
    best, bd = -1e300, 0
    for d in range(0, max_delay):
        c = float(np.dot(a[seg], b[seg.start + d:seg.stop + d]))
        if c > best:
            best, bd = c, d
    m = min(orig.shape[1], dec.shape[1] - bd)
    return bd, snr_db(orig[:, 2304:m].astype(np.float64), dec[:, bd + 2304:bd + m])
