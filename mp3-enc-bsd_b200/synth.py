"""Synthetic PCM for the BASELINE.json configs. All int16, planar [n_ch][n_samples].

numpy `default_rng(seed)`; deterministic, so the oracle/reference and the CUDA path see identical PCM.
"""
import numpy as np


def _to_i16(x):
    return np.clip(np.rint(x * 32767.0), -32768, 32767).astype(np.int16)


def music_channel(n, fs, seed):
    """Config-1 recipe: 0.25 sin 440 Hz + 0.15 FM tone (1 kHz +/- 500 Hz @0.3 Hz) + 0.05 N(0,1)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / fs
    fm_phase = 2 * np.pi * 1000.0 * t - (500.0 / 0.3) * np.cos(2 * np.pi * 0.3 * t)
    x = 0.25 * np.sin(2 * np.pi * 440.0 * t) + 0.15 * np.sin(fm_phase) + 0.05 * rng.standard_normal(n)
    return x


def config1(seconds=30.0, fs=44100, seeds=(1, 2)):
    """44.1 kHz stereo 'music' (configs 1, 4, 5; config 4 uses seeds=(2*clip+1, 2*clip+2))."""
    n = int(round(seconds * fs))
    return _to_i16(np.stack([music_channel(n, fs, s) for s in seeds]))


def config2(seconds=30.0, fs=32000, seed=3):
    """32 kHz mono transient-heavy: near silence + 20 ms full-scale decaying noise bursts every 250 ms
    + occasional 5 kHz tone bursts (drives pe > 1800 -> short blocks)."""
    rng = np.random.default_rng(seed)
    n = int(round(seconds * fs))
    x = 1e-3 * rng.standard_normal(n)
    burst = int(0.020 * fs)
    env = np.exp(-np.arange(burst) / (0.004 * fs))
    for k, start in enumerate(range(int(0.1 * fs), n - burst, int(0.25 * fs))):
        if k % 3 == 2:
            x[start:start + burst] += 0.8 * env * np.sin(2 * np.pi * 5000.0 * np.arange(burst) / fs)
        else:
            x[start:start + burst] += 0.95 * env * rng.uniform(-1, 1, burst)
    return _to_i16(x[None, :])


def config3(seconds=30.0, fs=48000, seeds=(4, 5)):
    """48 kHz stereo music-like: 40 random-phase partials 50 Hz-18 kHz with 1/f envelope + 0.1 pink-ish
    noise, L/R 0.7 correlated."""
    n = int(round(seconds * fs))
    t = np.arange(n, dtype=np.float64) / fs
    chans = []
    for s in seeds:
        rng = np.random.default_rng(s)
        f = np.exp(rng.uniform(np.log(50.0), np.log(18000.0), 40))
        ph = rng.uniform(0, 2 * np.pi, 40)
        x = np.zeros(n)
        for fi, pi_ in zip(f, ph):
            x += (50.0 / fi) ** 0.5 * np.sin(2 * np.pi * fi * t + pi_)
        white = rng.standard_normal(n)
        pink = np.cumsum(white) * 0.02
        pink -= np.convolve(pink, np.ones(512) / 512, mode="same")
        x = 0.6 * x / np.max(np.abs(x)) + 0.1 * pink / (np.max(np.abs(pink)) + 1e-12)
        chans.append(x)
    l, r = chans
    r = 0.7 * l + (1 - 0.7 ** 2) ** 0.5 * r
    m = max(np.max(np.abs(l)), np.max(np.abs(r)))
    return _to_i16(np.stack([l, r]) * (0.9 / m))


def clip_batch(n_clips, seconds=10.0, fs=44100, first=0):
    """Config 4: n_clips x stereo clips, config-1 recipe with seed = clip index. [n_clips][2][n]."""
    return np.stack([config1(seconds, fs, (2 * (first + c) + 1, 2 * (first + c) + 2)) for c in range(n_clips)])


def loud_sweep(seconds=1.0, fs=44100, seed=7):
    """Full-scale stereo test signal: loud low-frequency tone + sweep + clipped noise bursts. Drives
    |xr| >= 1 (so calc_scfsi's int-typed statistics become non-trivial) and large ix / ESC tables."""
    rng = np.random.default_rng(seed)
    n = int(round(seconds * fs))
    t = np.arange(n, dtype=np.float64) / fs
    f = 60.0 + (8000.0 - 60.0) * t / max(seconds, 1e-9)
    sweep = np.sin(2 * np.pi * np.cumsum(f) / fs)
    l = 0.98 * np.sin(2 * np.pi * 110.0 * t) * 0.6 + 0.5 * sweep
    r = 0.98 * np.sin(2 * np.pi * 110.0 * t + 0.3) * 0.6 + 0.45 * sweep
    burst = (np.floor(t / 0.1) % 4 == 3)
    l = np.where(burst, l + 0.9 * rng.uniform(-1, 1, n), l)
    r = np.where(burst, r + 0.9 * rng.uniform(-1, 1, n), r)
    return _to_i16(np.stack([l, r]))


def full_scale_tone(seconds=1.0, fs=44100, freq=1000.0, n_ch=2):
    """Full-scale sine on every channel: |xr| >= 1, so the reference's int-typed calc_scfsi statistics
    (loop.c:615-722) become non-zero and scfsi gets set; also saturates pow_nint at 2047 in the bin search."""
    n = int(round(seconds * fs))
    t = np.arange(n, dtype=np.float64) / fs
    return _to_i16(np.stack([np.sin(2 * np.pi * freq * t)] * n_ch))


# ---- heterogeneous clip batch for the throughput bench (BASELINE configs[3]) --------------------------------------
# torch ops only, so the same code generates the batch on the GPU (bench) and single clips on the CPU (reference arm,
# tests).  Every clip is a pure function of its GLOBAL index (counter-based noise, per-clip parameters from a hash), so a
# rank's shard holds the same clips whatever the world size is.
HETERO_CLASSES = ["tone+fm+noise", "tone+fm+noise", "partials", "transients", "loud-noise", "silence-then-music", "quiet-partials",
                  "am-tone"]


def _hash_u01(torch, a, b):
    """counter-based uniform [0,1): murmur-style finaliser of two int64 tensors (broadcast)"""
    x = (a * 0x9E3779B1 + b * 0x85EBCA6B + 0x165667B1) & 0xFFFFFFFF
    x = ((x ^ (x >> 16)) * 0x85EBCA6B) & 0xFFFFFFFF
    x = ((x ^ (x >> 13)) * 0xC2B2AE35) & 0xFFFFFFFF
    x = x ^ (x >> 16)
    return x.to(torch.float32) * (1.0 / 4294967296.0)


def hetero_batch(torch, first, count, n_samples, fs=44100, n_ch=2, device="cpu", sub=32, out=None, force_class=None, start=0):
    """int16 [count][n_ch][n_samples]: clips first .. first+count-1 of the heterogeneous bench batch.
    Class = index % 8 (HETERO_CLASSES): the config-1 recipe at levels spread over 30 dB and detuned, sums of partials
    (config-3 like), decaying noise / tone bursts on near silence (config-2 like, drives short blocks), loud white noise
    (bit pressure), digital silence followed by music, a -40 dB clip, an amplitude-modulated tone (configs[4] recipe).
    force_class: every clip takes that class (the per-clip parameters still come from the clip index).
    start: absolute index of the first sample (a clip is a pure function of the absolute sample index too, so a long clip
    can be generated piecewise)."""
    if out is None:
        out = torch.empty((count, n_ch, n_samples), dtype=torch.int16, device=device)
    two_pi = 6.283185307179586
    idx = torch.arange(start, start + n_samples, dtype=torch.int64, device=device)
    t = idx.to(torch.float64) / fs
    t32 = t.to(torch.float32)
    chs = torch.arange(n_ch, dtype=torch.int64, device=device).view(1, n_ch, 1)
    fm_dev = ((500.0 / 0.3) * torch.cos(two_pi * 0.3 * t))                       # config-1 FM phase deviation
    for s0 in range(0, count, sub):
        b = min(sub, count - s0)
        clip = torch.arange(first + s0, first + s0 + b, dtype=torch.int64, device=device).view(b, 1, 1)
        cls = clip % 8 if force_class is None else torch.full_like(clip, int(force_class))
        par = lambda k: _hash_u01(torch, clip, torch.full_like(clip, 1000003 + k))          # per-clip parameter k in [0,1)
        # white noise ~ N(0,1): sum of four uniforms
        key = clip * 4 + chs
        u = sum(_hash_u01(torch, key * 7919 + j, idx.view(1, 1, -1)) for j in range(4))
        noise = (u - 2.0) * 1.7320508
        level = torch.pow(10.0, -1.5 * par(0))                                   # 0 .. -30 dB
        f0 = 440.0 * torch.pow(2.0, 2.0 * par(1) - 1.0)                          # 220 .. 880 Hz
        ph_tone = (two_pi * f0.to(torch.float64) * t.view(1, 1, -1))
        tone = torch.sin(ph_tone).to(torch.float32)
        fm = torch.sin(two_pi * 1000.0 * t - fm_dev).to(torch.float32).view(1, 1, -1)
        x_c1 = (0.25 * tone + 0.15 * fm + 0.05 * noise) * level
        # partials: 12 log-uniform frequencies 50 Hz .. 12 kHz, amplitude ~ 1 / sqrt(f), channel-dependent phases
        part = torch.zeros((b, n_ch, n_samples), dtype=torch.float32, device=device)
        norm = torch.zeros((b, 1, 1), dtype=torch.float32, device=device)
        for k in range(12):
            fk = 50.0 * torch.pow(240.0, par(10 + k))
            ak = torch.rsqrt(fk / 50.0)
            phk = two_pi * _hash_u01(torch, key, torch.full_like(key, 77 + k))
            part += ak * torch.sin((two_pi * fk.to(torch.float64) * t.view(1, 1, -1)).to(torch.float32) + phk)
            norm += ak
        x_part = 0.7 * part / norm + 0.01 * noise
        # transients: near silence + a 20 ms decaying burst every 250 ms (noise, every third one a 5 kHz tone)
        tb = torch.remainder(t32 - 0.1, 0.25)
        kb = torch.floor((t32 - 0.1) / 0.25)
        env = torch.where((t32 >= 0.1) & (tb < 0.020), torch.exp(-tb / 0.004), torch.zeros_like(tb)).view(1, 1, -1)
        tone5k = torch.sin(two_pi * 5000.0 * t).to(torch.float32).view(1, 1, -1)
        is_tone = (torch.remainder(kb, 3.0) == 2.0).view(1, 1, -1)
        x_tr = 1e-3 * noise + env * torch.where(is_tone, 0.8 * tone5k, 0.95 * 0.5773503 * noise)
        x_loud = 0.28 * noise                                                     # clipped rarely, every band full
        gate = (t32 >= 5.0).view(1, 1, -1)
        x_sil = torch.where(gate, x_c1 / level * 0.7, torch.zeros_like(x_c1))     # digital silence, then the recipe at -3 dB
        x_quiet = 0.01 * x_part
        am = (0.6 + 0.4 * torch.sin(two_pi * 0.05 * t + two_pi * par(2).to(torch.float64))).to(torch.float32)
        x_am = (0.25 * tone + 0.15 * fm + 0.05 * noise) * am
        x = torch.where(cls <= 1, x_c1, torch.where(cls == 2, x_part, torch.where(cls == 3, x_tr, torch.where(
            cls == 4, x_loud, torch.where(cls == 5, x_sil, torch.where(cls == 6, x_quiet, x_am))))))
        out[s0:s0 + b] = torch.clamp(torch.round(x * 32767.0), -32768, 32767).to(torch.int16)
    return out


def hetero_stream(torch, start, count, fs=44100, n_ch=2, device="cpu"):
    """samples [start, start+count) of THE long stream of BASELINE configs[4] (clip 0 as class "am-tone": the config-1
    recipe under a slow amplitude envelope): int16 [n_ch][count]"""
    return hetero_batch(torch, 0, 1, count, fs, n_ch, device, force_class=7, start=start)[0]


# ---- machine-independent clips (IEEE add / multiply and integer arithmetic only) -------------------------------------------
# numpy's sin / exp may differ in the last bit between CPUs (SIMD code paths), which after rounding to int16 can change a
# sample; fixtures that are stored as HASHES of the reference's output need PCM that is the same everywhere.
def _lcg_noise(n, seed):
    """uniform noise in [-1, 1): 64-bit LCG (Knuth's MMIX constants) evaluated by jumping, top 24 bits"""
    a, c, m = 6364136223846793005, 1442695040888963407, (1 << 64) - 1
    # x_k = a^k x_0 + c (a^k - 1) / (a - 1): evaluate blockwise with Python big ints for the block heads, numpy inside
    out = np.empty(n, np.float64)
    x = (seed * 2654435761 + 12345) & m
    block = 1 << 16
    ak, ck = np.uint64(a), np.uint64(c)
    for b0 in range(0, n, block):
        k = min(block, n - b0)
        v = np.empty(k, np.uint64)
        cur = np.uint64(x)
        with np.errstate(over="ignore"):
            for i in range(k):
                cur = cur * ak + ck
                v[i] = cur
        x = int(cur)
        out[b0:b0 + k] = (v >> np.uint64(40)).astype(np.float64) / float(1 << 23) - 1.0
    return out


def _parabolic_tone(n, fs, freq_mhz, phase0=0):
    """sine-like wave from integer phase arithmetic: p = frac(f t) exactly (integers), s = 16 p (1 - p) on each half wave"""
    t = np.arange(n, dtype=np.int64)
    period_units = fs * 1000                                  # phase in units of 1 / (fs * 1000) cycles
    ph = (t * freq_mhz + phase0) % period_units               # exact integers (freq in millihertz)
    p = ph.astype(np.float64) / float(period_units)           # one correctly rounded division
    half = p < 0.5
    q = np.where(half, p, p - 0.5) * 2.0
    s = 4.0 * q * (1.0 - q)
    return np.where(half, s, -s)


def exact_clip(kind, seconds, fs, n_ch, seed=1):
    """int16 [n_ch][n], identical on every IEEE machine.  kind: "music" (three tones + noise), "transient" (decaying noise
    bursts on near silence every 250 ms: drives short blocks), "loud" (full-band noise at -10 dB: bit pressure)"""
    n = int(round(seconds * fs))
    chans = []
    for ch in range(n_ch):
        noise = _lcg_noise(n, 1000 * seed + ch)
        if kind == "music":
            x = 0.25 * _parabolic_tone(n, fs, 440000 + 7000 * ch) + 0.15 * _parabolic_tone(n, fs, 1330500, 12345 * (ch + 1)) + \
                0.1 * _parabolic_tone(n, fs, 5200250) + 0.04 * noise
        elif kind == "transient":
            t = np.arange(n, dtype=np.int64)
            per, burst = fs // 4, fs // 50
            k = (t - fs // 10) % per
            on = (t >= fs // 10) & (k < burst)
            env = np.where(on, (1.0 - k.astype(np.float64) / float(burst)) ** 2, 0.0)
            env = env * env                                      # (1 - k / burst)^4: fast decay, arithmetic only
            x = 1e-3 * noise + 0.9 * env * noise
        elif kind == "loud":
            x = 0.3 * noise
        else:
            raise ValueError(kind)
        chans.append(x)
    return _to_i16(np.stack(chans))
