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
