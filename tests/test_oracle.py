"""CPU tests: the oracle (oracle/fm_oracle.c) against the golden vectors generated from the
unmodified reference, and -- where oracle/_ref is available -- against the compiled reference
itself on fresh inputs.  Bit-exact everywhere (uint32 views of float32)."""
import hashlib
import os

import numpy as np
import pytest

import orclib
import sdr_b200
from sdr_b200 import siggen

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TAPSETS = {"F": (151, 101, 151), "S": (13, 13, 13)}


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_libm_replicas_match_golden(orc):
    g = np.load(os.path.join(GOLD, "libm.npz"))
    x = g["x"]
    s = np.empty_like(x); c = np.empty_like(x)
    orc.lib.orc_sincosf_batch(x, x.size, s, c)
    assert np.array_equal(bits(s), bits(g["sin"]))
    assert np.array_equal(bits(c), bits(g["cos"]))
    assert np.array_equal(bits(c), bits(g["cosf"]))  # cosf == cos half of sincosf
    a = np.empty_like(g["ay"])
    orc.lib.orc_atan2f_batch(g["ay"], g["ax"], a.size, a)
    assert np.array_equal(bits(a), bits(g["atan2"]))


def test_libm_replicas_match_host_libm(orc, ref):
    """Pins the replicas against the libm of THIS host (the one _ref links to)."""
    rng = np.random.default_rng(11)
    n = 400_000
    x = np.concatenate([rng.uniform(-4, 4, n), rng.uniform(-200, 200, n), rng.uniform(-3e5, 3e5, n),
                        rng.standard_normal(n) * 10.0 ** rng.uniform(-20, 20, n)]).astype(np.float32)
    s1 = np.empty_like(x); c1 = np.empty_like(x); s2 = np.empty_like(x); c2 = np.empty_like(x)
    c3 = np.empty_like(x)
    orc.lib.orc_sincosf_batch(x, x.size, s1, c1)
    ref.lib.ref_libm_sincosf(x, x.size, s2, c2)
    ref.lib.ref_libm_cosf(x, x.size, c3)
    assert np.array_equal(bits(s1), bits(s2)) and np.array_equal(bits(c1), bits(c2))
    assert np.array_equal(bits(c1), bits(c3))
    y = (rng.standard_normal(4 * n) * 10.0 ** rng.uniform(-6, 2, 4 * n)).astype(np.float32)
    xx = (rng.standard_normal(4 * n) * 10.0 ** rng.uniform(-6, 2, 4 * n)).astype(np.float32)
    a1 = np.empty_like(y); a2 = np.empty_like(y)
    orc.lib.orc_atan2f_batch(y, xx, y.size, a1)
    ref.lib.ref_libm_atan2f(y, xx, y.size, a2)
    assert np.array_equal(bits(a1), bits(a2))


def test_design_matches_golden(orc):
    g = np.load(os.path.join(GOLD, "design.npz"))
    for k in g.files:
        if k.endswith("_args"):
            continue
        args = g[k + "_args"]
        if k.startswith("lpf_"):
            got = orc.lpf(float(args[0]), float(args[1]), int(args[2]))
            mine = sdr_b200.impulseResponseLPF(float(args[0]), float(args[1]), int(args[2]))
        else:
            got = orc.bpf(float(args[0]), float(args[1]), float(args[2]), int(args[3]))
            mine = sdr_b200.bandPass(float(args[0]), float(args[1]), float(args[2]), int(args[3]))
        assert np.array_equal(bits(got), bits(g[k])), k
        # the product's host-side design (design.cpp) is held to the same vectors
        assert np.array_equal(bits(mine), bits(g[k])), k


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_chain_matches_golden(orc, mode):
    g = np.load(os.path.join(GOLD, f"chain_mode{mode}.npz"))
    iq = g["iq"]
    for ch in (1, 2):
        for tname, taps in TAPSETS.items():
            pcm, t = orc.run_chain(iq, mode, ch, *taps)
            key = f"c{ch}_{tname}"
            assert np.array_equal(pcm, g[key + "_pcm"]), key
            for name, arr in t.items():
                assert int(g[f"{key}_n_{name}"]) == arr.size, (key, name)
                assert str(g[f"{key}_sha_{name}"]) == sha(arr), (key, name)


def test_chain_matches_golden_on_recorded_iq(orc):
    """The only recorded I/Q data the reference ships (two blocks its front end printed into
    data/data/pipeData.txt; tests/golden/make_golden_real.py recovers the bytes) and what the
    unmodified reference makes of it."""
    g = np.load(os.path.join(GOLD, "real_iq.npz"))
    for ch in (1, 2):
        for tname, taps in TAPSETS.items():
            pcm, t = orc.run_chain(g["iq"], 0, ch, *taps)
            key = f"c{ch}_{tname}"
            assert np.array_equal(pcm, g[key + "_pcm"]), key
            for name, arr in t.items():
                assert str(g[f"{key}_sha_{name}"]) == sha(arr), (key, name)


@pytest.mark.parametrize("mode,ch", [(0, 1), (0, 2), (1, 2), (2, 1), (2, 2), (3, 2)])
def test_chain_matches_compiled_reference(orc, ref, mode, ch):
    iq = siggen.make_capture(7 * mode + ch, mode, 3, "stereo")
    for taps in TAPSETS.values():
        p1, t1 = orc.run_chain(iq, mode, ch, *taps)
        p2, t2 = ref.run_chain(iq, mode, ch, *taps)
        assert np.array_equal(p1, p2)
        assert set(t1) == set(t2)
        for k in t2:
            assert np.array_equal(bits(t1[k]), bits(t2[k])), k


def test_primitives_match_compiled_reference(orc, ref):
    rng = np.random.default_rng(5)
    x = rng.standard_normal(3000).astype(np.float32)
    h = rng.standard_normal(37).astype(np.float32)
    for decim in (1, 3, 5):
        s1 = rng.standard_normal(36).astype(np.float32); s2 = s1.copy()
        y1 = np.zeros(3000 // decim, np.float32); y2 = y1.copy()
        orc.fir_decim(y1, x, x.size, h, h.size, s1, decim)
        ref.fir_decim(y2, x, x.size, h, h.size, s2, decim)
        assert np.array_equal(bits(y1), bits(y2)) and np.array_equal(bits(s1), bits(s2))
    s1 = rng.standard_normal(36).astype(np.float32); s2 = s1.copy()
    y1 = np.zeros(3000, np.float32); y2 = y1.copy()
    orc.fir_block(y1, x, x.size, h, h.size, s1)
    ref.fir_block(y2, x, x.size, h, h.size, s2)
    assert np.array_equal(bits(y1), bits(y2)) and np.array_equal(bits(s1), bits(s2))
    # the small resampler shapes of the reference's scratch harness (src/testing.cpp:62-65,75-92)
    for U, D, tp, nx in ((3, 4, 17, 36), (147, 800, 13, 5600), (7, 5, 9, 700)):
        hh = rng.standard_normal(tp * U).astype(np.float32)
        xx = rng.standard_normal(nx).astype(np.float32)
        s1 = np.zeros(tp * U - 1, np.float32); s2 = s1.copy()
        for _ in range(2):  # second pass exercises the carried zero-stuffed state
            y1 = np.zeros(nx * U // D, np.float32); y2 = y1.copy()
            orc.fir_resample(y1, xx, nx, hh, hh.size, s1, D, U)
            ref.fir_resample(y2, xx, nx, hh, hh.size, s2, D, U)
            assert np.array_equal(bits(y1), bits(y2)) and np.array_equal(bits(s1), bits(s2))
    raw = rng.integers(0, 256, 4096, dtype=np.uint8)
    f1 = np.zeros(4096, np.float32); f2 = f1.copy()
    orc.u8_to_f32(raw, raw.size, f1); ref.u8_to_f32(raw, raw.size, f2)
    assert np.array_equal(bits(f1), bits(f2))
    for v in (0.0, 0.5, -0.5, 1.99999, -2.0, 2.0, 3.7, -1e9, 1e9, 131071.9, float("nan"), float("inf")):
        assert orc.pcm16(v) == ref.pcm16(v), v


def test_silence_hits_zero_denominator_branch(orc):
    """All-128 input: I=Q=0 -> fmDemod's `== 0` branch (filter.cpp:254) -> PCM all zero."""
    iq = siggen.make_capture(0, 0, 2, "silence")
    pcm, t = orc.run_chain(iq, 0, 2)
    assert not pcm.any() and not t["demod"].any()


def test_siggen_is_deterministic_and_in_format():
    a = siggen.make_capture(3, 0, 1, "stereo")
    b = siggen.make_capture(3, 0, 1, "stereo")
    assert a.dtype == np.uint8 and a.size == siggen.MODES[0]["block_bytes"]
    assert np.array_equal(a, b)
    assert not np.array_equal(a, siggen.make_capture(4, 0, 1, "stereo"))
    batch = siggen.make_batch(5, 1, 1, "mono", distinct=2)
    assert batch.shape == (5, siggen.MODES[1]["block_bytes"])
    assert not np.array_equal(batch[0], batch[2])
