"""Synthetic FM broadcast captures (host tooling; numpy only).

The reference ships no I/Q captures (its drivers read ``../data/lab3_iq_samples/*.raw``
which is absent, model/fmMonoBasic.py:91), so every parity and throughput run uses the
deterministic generator below.  The output format is the one the reference consumes:
interleaved unsigned 8-bit I,Q (``rtl_sdr`` format, src/iofunc.cpp:128-135).

Channel ``c`` is seeded with ``0x5D2 + c`` so captures are reproducible on any host.
"""
from __future__ import annotations

import numpy as np

# (rf_Fs, rf_decim, if_Fs, audio_decim, audio_upsamp, audio_Fs, block_bytes) per mode,
# src/project.cpp:424-427 and :55-57.
MODES = {
    0: dict(rf_Fs=2_400_000, rf_decim=10, if_Fs=240_000, audio_decim=5, audio_upsamp=1,
            audio_Fs=48_000, block_bytes=1024 * 10 * 5 * 2),
    1: dict(rf_Fs=1_440_000, rf_decim=5, if_Fs=288_000, audio_decim=6, audio_upsamp=1,
            audio_Fs=48_000, block_bytes=1024 * 5 * 6 * 2),
    2: dict(rf_Fs=2_400_000, rf_decim=10, if_Fs=240_000, audio_decim=800, audio_upsamp=147,
            audio_Fs=44_100, block_bytes=7 * 800 * 10 * 2),
    3: dict(rf_Fs=960_000, rf_decim=3, if_Fs=320_000, audio_decim=3200, audio_upsamp=441,
            audio_Fs=44_100, block_bytes=7 * 3200 * 3 * 2),
}

SEED_BASE = 0x5D2


def _audio(rng: np.random.Generator, t: np.ndarray) -> np.ndarray:
    """Three seed-drawn tones in 200 Hz..14 kHz plus a little low-passed noise, |x| <= 0.45."""
    f = rng.uniform(200.0, 14_000.0, size=3)
    a = rng.uniform(0.05, 0.14, size=3)
    ph = rng.uniform(0, 2 * np.pi, size=3)
    x = sum(a[i] * np.sin(2 * np.pi * f[i] * t + ph[i]) for i in range(3))
    noise = rng.standard_normal(t.size)
    # crude band limit: moving average over ~1/12 kHz
    k = max(1, int(round(1.0 / (12_000.0 * (t[1] - t[0])))))
    noise = np.convolve(noise, np.ones(k) / k, mode="same") * 0.02
    return np.clip(x + noise, -0.45, 0.45)


def _rds_baseband(rng: np.random.Generator, t: np.ndarray) -> np.ndarray:
    """Differentially encoded, Manchester (biphase) coded bit stream at 1187.5 bit/s,
    half-sine shaped.  Only used to put energy at 57 kHz for the RDS rows."""
    bitrate = 1187.5
    nbits = int(np.ceil(t[-1] * bitrate)) + 2
    bits = rng.integers(0, 2, size=nbits)
    diff = np.bitwise_xor.accumulate(bits)
    chips = np.empty(2 * nbits)
    chips[0::2] = 2.0 * diff - 1.0
    chips[1::2] = -(2.0 * diff - 1.0)
    pos = t * (2 * bitrate)
    idx = np.minimum(pos.astype(np.int64), chips.size - 1)
    return chips[idx] * np.sin(np.pi * (pos - idx))


def make_capture(channel: int, mode: int, n_blocks: int, kind: str = "stereo",
                 cnr_db: float | None = None) -> np.ndarray:
    """Return ``n_blocks`` reference blocks of interleaved uint8 I/Q for one channel.

    kind: "mono" (L+R only), "stereo" (pilot + 38 kHz DSB-SC), "rds" (stereo + 57 kHz RDS),
          "silence" (all bytes 128: exercises fmDemod's zero-denominator branch,
          src/filter.cpp:254), "clipped" (over-driven, saturating the 8-bit range).
    """
    m = MODES[mode]
    n = n_blocks * m["block_bytes"] // 2
    if kind == "silence":
        return np.full(2 * n, 128, dtype=np.uint8)
    rng = np.random.default_rng(SEED_BASE + channel)
    fs = float(m["rf_Fs"])
    t = np.arange(n, dtype=np.float64) / fs
    left = _audio(rng, t)
    right = _audio(rng, t)
    mpx = 0.45 * (left + right)
    if kind in ("stereo", "rds", "clipped"):
        mpx = mpx + 0.1 * np.sin(2 * np.pi * 19_000.0 * t)
        mpx = mpx + 0.45 * (left - right) * np.sin(2 * np.pi * 38_000.0 * t)
    if kind == "rds":
        mpx = mpx + 0.05 * _rds_baseband(rng, t) * np.sin(2 * np.pi * 57_000.0 * t)
    f_off = rng.uniform(-5_000.0, 5_000.0)
    phase = 2 * np.pi * np.cumsum(75_000.0 * mpx + f_off) / fs
    amp = 1.6 if kind == "clipped" else 0.6
    if cnr_db is None:
        cnr_db = rng.uniform(30.0, 50.0)
    sigma = amp * 10.0 ** (-cnr_db / 20.0) / np.sqrt(2.0)
    i = amp * np.cos(phase) + sigma * rng.standard_normal(n)
    q = amp * np.sin(phase) + sigma * rng.standard_normal(n)
    out = np.empty(2 * n, dtype=np.uint8)
    out[0::2] = np.clip(np.rint(127.5 + 127.5 * i), 0, 255).astype(np.uint8)
    out[1::2] = np.clip(np.rint(127.5 + 127.5 * q), 0, 255).astype(np.uint8)
    return out


def make_batch(n_channels: int, mode: int, n_blocks: int, kind: str = "stereo",
               distinct: int | None = None) -> np.ndarray:
    """``[n_channels, nbytes]`` batch.  With ``distinct`` < n_channels the first ``distinct``
    captures are generated and then repeated with a per-channel circular byte-pair rotation,
    which keeps every channel a valid (different) FM capture while bounding generation time."""
    distinct = n_channels if distinct is None else min(distinct, n_channels)
    base = [make_capture(c, mode, n_blocks, kind) for c in range(distinct)]
    out = np.empty((n_channels, base[0].size), dtype=np.uint8)
    for c in range(n_channels):
        src = base[c % distinct]
        rot = 2 * ((c // distinct) * 977 % (src.size // 2))
        out[c] = np.roll(src, rot) if rot else src
    return out
