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


# ---- RDS with a known group sequence ---------------------------------------------------------
# Parity-check matrix and offset-word syndromes of the RDS standard as the reference's model
# holds them (model/fmSupportLib.py:32-57, :62-91).
_RDS_H = np.array([
    [1,0,0,0,0,0,0,0,0,0], [0,1,0,0,0,0,0,0,0,0], [0,0,1,0,0,0,0,0,0,0], [0,0,0,1,0,0,0,0,0,0],
    [0,0,0,0,1,0,0,0,0,0], [0,0,0,0,0,1,0,0,0,0], [0,0,0,0,0,0,1,0,0,0], [0,0,0,0,0,0,0,1,0,0],
    [0,0,0,0,0,0,0,0,1,0], [0,0,0,0,0,0,0,0,0,1], [1,0,1,1,0,1,1,1,0,0], [0,1,0,1,1,0,1,1,1,0],
    [0,0,1,0,1,1,0,1,1,1], [1,0,1,0,0,0,0,1,1,1], [1,1,1,0,0,1,1,1,1,1], [1,1,0,0,0,1,0,0,1,1],
    [1,1,0,1,0,1,0,1,0,1], [1,1,0,1,1,1,0,1,1,0], [0,1,1,0,1,1,1,0,1,1], [1,0,0,0,0,0,0,0,0,1],
    [1,1,1,1,0,1,1,1,0,0], [0,1,1,1,1,0,1,1,1,0], [0,0,1,1,1,1,0,1,1,1], [1,0,1,0,1,0,0,1,1,1],
    [1,1,1,0,0,0,1,1,1,1], [1,1,0,0,0,1,1,0,1,1]], dtype=np.int64)
_RDS_SYNDROME = {"A": [1,1,1,1,0,1,1,0,0,0], "B": [1,1,1,1,0,1,0,1,0,0], "C": [1,0,0,1,0,1,1,1,0,0],
                 "D": [1,0,0,1,0,1,1,0,0,0]}
_RDS_CHECKS = np.array([[(c >> (9 - i)) & 1 for i in range(10)] for c in range(1024)], dtype=np.int64)
_RDS_CHECK_SYN = _RDS_CHECKS @ _RDS_H[16:] % 2
# Delay of the chip clock that puts the chip centres, after every filter of the receiver, on the
# model's fixed sampling grid 158 + i*SPS (model/fmRDS.py:259): measured with the oracle.
_RDS_CHIP_DELAY = {0: 2.0 / 61_750.0, 2: 19.0 / 102_125.0}


def rds_block(info16: np.ndarray, offset: str) -> np.ndarray:
    """26-bit RDS block: 16 information bits + the 10 check bits that give the block the
    syndrome of `offset` ('A', 'B', 'C' or 'D') under the model's parity-check matrix."""
    need = (np.array(_RDS_SYNDROME[offset]) + np.asarray(info16, np.int64) @ _RDS_H[:16]) % 2
    idx = np.where((_RDS_CHECK_SYN == need).all(axis=1))[0]
    assert idx.size == 1
    return np.concatenate((np.asarray(info16, np.int64), _RDS_CHECKS[idx[0]]))


def rds_group_bits(channel: int, mode: int, n_blocks: int) -> np.ndarray:
    """The bit sequence (before differential encoding) that kind="rds_groups" transmits: groups
    of four valid blocks A, B, C, D with seed-drawn information words."""
    m = MODES[mode]
    seconds = n_blocks * m["block_bytes"] // 2 / float(m["rf_Fs"])
    nbits = int(seconds * 1187.5) + 64
    rng = np.random.default_rng(SEED_BASE + 7919 * (channel + 1))
    blocks = []
    while 26 * len(blocks) < nbits:
        for off in "ABCD":
            blocks.append(rds_block(rng.integers(0, 2, 16), off))
    return np.concatenate(blocks)[:nbits].astype(np.uint8)


def _rrc_pulse(t: np.ndarray, T: float, beta: float = 0.9) -> np.ndarray:
    x = t / T
    with np.errstate(divide="ignore", invalid="ignore"):
        out = (np.sin(np.pi * x * (1 - beta)) + 4 * beta * x * np.cos(np.pi * x * (1 + beta))) / \
              (np.pi * x * (1 - (4 * beta * x) ** 2))
    out[np.abs(x) < 1e-9] = 1 - beta + 4 * beta / np.pi
    out[np.abs(np.abs(x) - 1 / (4 * beta)) < 1e-9] = beta / np.sqrt(2) * (
        (1 + 2 / np.pi) * np.sin(np.pi / (4 * beta)) + (1 - 2 / np.pi) * np.cos(np.pi / (4 * beta)))
    return out


def _rds_group_baseband(bits: np.ndarray, t: np.ndarray, delay: float) -> np.ndarray:
    """Differential encoding, biphase chips at 2375 chips/s (bit 1 = high, low), each chip a
    root-raised-cosine pulse (roll-off 0.9, the receiver's matched filter)."""
    d = np.bitwise_xor.accumulate(bits.astype(np.int64))
    chips = np.empty(2 * d.size)
    chips[0::2] = 2.0 * d - 1.0
    chips[1::2] = -(2.0 * d - 1.0)
    T = 1.0 / 2375.0
    n0 = np.floor((t - delay) / T).astype(np.int64)
    s = np.zeros_like(t)
    for k in range(-3, 5):
        n = n0 + k
        ok = (n >= 0) & (n < chips.size)
        s += np.where(ok, chips[np.clip(n, 0, chips.size - 1)], 0.0) * _rrc_pulse(t - delay - n * T, T)
    return s


def make_capture(channel: int, mode: int, n_blocks: int, kind: str = "stereo",
                 cnr_db: float | None = None) -> np.ndarray:
    """Return ``n_blocks`` reference blocks of interleaved uint8 I/Q for one channel.

    kind: "mono" (L+R only), "stereo" (pilot + 38 kHz DSB-SC), "rds" (stereo + 57 kHz RDS),
          "rds_groups" (stereo + an RDS subcarrier that carries rds_group_bits(): valid groups
          of blocks A-D, root-raised-cosine chips timed for the model's sampling grid; modes 0, 2),
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
    if kind == "rds_groups":
        bits = rds_group_bits(channel, mode, n_blocks)
        mpx = 0.4 * (left + right) + 0.1 * np.sin(2 * np.pi * 19_000.0 * t)
        mpx = mpx + 0.4 * (left - right) * np.sin(2 * np.pi * 38_000.0 * t)
        mpx = mpx + 0.08 * _rds_group_baseband(bits, t, _RDS_CHIP_DELAY[mode]) * \
            np.cos(2 * np.pi * 57_000.0 * t + 0.3)
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


def make_wideband(n_channels: int, n_pairs_per_channel: int, tone_hz=None, fs_channel: float = 2.4e6,
                  skip=()) -> tuple[np.ndarray, float]:
    """One wideband capture (interleaved uint8 I/Q at n_channels * fs_channel) holding one mono FM
    station per channel of an n_channels-band channeliser: station c sits c * fs_channel above the
    centre (c >= n_channels/2: below it, the DFT's wrap) and carries a single tone of tone_hz[c]
    (default 1000 + 400 c Hz) at 37.5 kHz deviation.  Returns (bytes, amplitude of one station)."""
    M = n_channels
    n = n_pairs_per_channel * M
    fs = fs_channel * M
    t = np.arange(n, dtype=np.float64) / fs
    amp = 0.8 / M
    x = np.zeros(n, np.complex128)
    for c in range(M):
        if c in skip:
            continue
        f_tone = (1000.0 + 400.0 * c) if tone_hz is None else tone_hz[c]
        f_c = (c if c < M / 2 else c - M) * fs_channel
        # phase of the carrier offset is exact per sample (f_c / fs is a multiple of 1/M)
        dev = 37_500.0 / (2 * np.pi * f_tone) * (-np.cos(2 * np.pi * f_tone * t) + 1.0) * 2 * np.pi
        k = (c if c < M / 2 else c - M)
        x += amp * np.exp(1j * (dev + 2 * np.pi * ((k * np.arange(n)) % M) / M))
    out = np.empty(2 * n, np.uint8)
    out[0::2] = np.clip(np.rint(128 + 128 * x.real), 0, 255).astype(np.uint8)
    out[1::2] = np.clip(np.rint(128 + 128 * x.imag), 0, 255).astype(np.uint8)
    return out, amp
