"""CPU tests of the checkers for the stages either side of the hot path (oracle/aux_oracle.c):
PSD restatement pinned to the compiled reference and to golden vectors made from it; the
de-emphasis and channeliser definitions (no reference implementation: parity unpinned) checked
through the properties that define them.  Also the WAV header (host-only code of the product)."""
import os
import struct

import numpy as np
import pytest

import auxlib
import orclib
import sdr_b200
from sdr_b200 import siggen

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_psd_oracle_matches_golden_vectors_of_the_reference():
    g = np.load(os.path.join(GOLD, "psd.npz"))
    for name in ("audio48k", "if240k", "one_segment"):
        freq, psd = auxlib.psd(g[f"{name}_x"], float(g[f"{name}_fs"]))
        assert np.array_equal(freq, g[f"{name}_freq"]), name
        assert float(np.abs(psd - g[f"{name}_psd"]).max()) <= 1e-4, name   # dB


def test_psd_oracle_matches_compiled_reference(ref):
    rng = np.random.default_rng(3)
    x = (0.2 * rng.standard_normal(512 * 4 + 17)).astype(np.float32)
    fo, po = auxlib.psd(x, 240000.0)
    fr, pr = auxlib.ref_psd(x, 240000.0)
    assert np.array_equal(fo, fr)
    assert float(np.abs(po - pr).max()) <= 1e-4


def test_deemphasis_oracle_is_a_75us_one_pole():
    """-3 dB at 1/(2 pi tau) = 2122 Hz, unity gain at DC, state carried across calls."""
    fs, tau = 48000.0, 75e-6
    t = np.arange(48000) / fs
    gains = {}
    for f in (50.0, 2122.0, 10000.0):
        pcm = np.rint(8000 * np.sin(2 * np.pi * f * t)).astype(np.int16)
        st = np.zeros(1, np.float32)
        y = auxlib.deemphasis(pcm.copy(), 1, fs, tau, st).astype(np.float64)
        gains[f] = np.sqrt(np.mean(y[4800:] ** 2)) / np.sqrt(np.mean(pcm[4800:].astype(np.float64) ** 2))
    assert abs(20 * np.log10(gains[50.0])) < 0.1
    assert abs(20 * np.log10(gains[2122.0]) + 3.0) < 0.35
    assert 20 * np.log10(gains[10000.0]) < -12.0
    pcm = np.rint(8000 * np.sin(2 * np.pi * 700.0 * t[:4000])).astype(np.int16)
    st = np.zeros(1, np.float32)
    whole = auxlib.deemphasis(pcm.copy(), 1, fs, tau, st)
    st = np.zeros(1, np.float32)
    parts = np.concatenate([auxlib.deemphasis(pcm[:1234].copy(), 1, fs, tau, st),
                            auxlib.deemphasis(pcm[1234:].copy(), 1, fs, tau, st)])
    assert np.array_equal(whole, parts)


def test_channelizer_oracle_separates_stations():
    """Eight stations, one per band: every output row holds its own station (a constant-envelope
    FM signal of the expected amplitude) and the skipped band is empty."""
    M, T = 8, 12
    wide, amp = siggen.make_wideband(M, 3000, skip=(3,))
    h = sdr_b200.impulseResponseLPF(float(M), 0.4, M * T)
    out = auxlib.channelize(wide, M, h, 0.6 / amp).astype(np.float64)
    for c in range(M):
        z = (out[c, 0::2] - 128) + 1j * (out[c, 1::2] - 128)
        env = np.abs(z[200:])
        if c == 3:
            assert env.max() < 6
        else:
            assert abs(env.mean() - 0.6 * 128) < 4 and env.std() < 4, (c, env.mean(), env.std())


def test_wav_header():
    h = sdr_b200.wav_header(48000, 2, 1000)
    assert len(h) == 44 and h[:4] == b"RIFF" and h[8:16] == b"WAVEfmt " and h[36:40] == b"data"
    riff, = struct.unpack("<I", h[4:8])
    fmt, ch, rate, brate, align, bits = struct.unpack("<HHIIHH", h[20:36])
    data, = struct.unpack("<I", h[40:44])
    assert (fmt, ch, rate, brate, align, bits) == (1, 2, 48000, 192000, 4, 16)
    assert data == 4000 and riff == 36 + 4000
    with pytest.raises(sdr_b200.SdrError):
        sdr_b200.wav_header(48000, 3, 10)
