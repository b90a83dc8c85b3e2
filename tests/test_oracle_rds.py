"""RDS oracle (oracle/rds_oracle.c) against the golden vectors that
tests/golden/make_golden_rds.py produced by importing the reference's Python model
(model/fmSupportLib.py + scipy.signal.lfilter, the calls of model/fmRDS.py:222-276)."""
import os

import numpy as np
import pytest

import orclib
from sdr_b200 import siggen

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FTOL = 1e-12  # of full scale: numpy/scipy sum the same products in another order


def fm_demod_of(mode, n_ref_blocks):
    """fm_demod of the golden capture from the C oracle (bit-identical to the reference's C++
    front end, tests/test_oracle.py)."""
    iq = siggen.make_capture(200, mode, n_ref_blocks, "rds")
    _, taps = orclib.ORC().run_chain(iq, mode, 1)
    return taps["demod"].astype(np.float64)


def close(got, want, tol=FTOL):
    scale = max(1.0, float(np.abs(want).max()))
    assert got.shape == want.shape
    assert float(np.abs(got - want).max()) <= tol * scale


def test_design_matches_model():
    R = orclib.RDS()
    g = np.load(os.path.join(GOLD, "rds_design.npz"))
    for mode, U, sps in ((0, 247, 26), (2, 817, 43)):
        sets = dict(chan=R.bandpass(151, 240000.0, 54e3, 60e3),
                    carr=R.bandpass(151, 240000.0, 113.5e3, 114.5e3),
                    rs=R.lowpass(101 * U, 240000.0 * U, 3e3),
                    rrc=R.rrc(2375.0 * sps, 101))
        for k, h in sets.items():
            want = g[f"m{mode}_{k}"]
            close(h if h.size <= 151 else h[::97], want, 1e-14)
            s = g[f"m{mode}_{k}_sum"]
            assert abs(h.sum() - s[0]) <= 1e-11 * s[1]


@pytest.mark.parametrize("mode,n_ref,block_if,n_blocks", [(0, 15, 9600, 8), (2, 12, 19200, 2)])
def test_chain_matches_model(mode, n_ref, block_if, n_blocks):
    g = np.load(os.path.join(GOLD, f"rds_mode{mode}.npz"))
    fm = fm_demod_of(mode, n_ref)[:block_if * n_blocks]
    keep = tuple(orclib.RDS_TAP_NAMES)
    r = orclib.RDS().run_chain(fm, mode, block_if, keep=keep)
    close(r["rrc_i"], g["rrc_i"])
    close(r["rrc_q"], g["rrc_q"])
    # every other stage, last block, every 16th sample
    # (the PLL turns the band-pass output's rounding differences, relative to that small
    # signal, into phase: 1e-11 rad here, so the NCO outputs get a wider absolute bound)
    for k in keep[:8]:
        per_block = r[k].size // n_blocks
        close(r[k][-per_block:][::16], g["last16_" + k], 1e-9 if k.startswith("pll") else FTOL)
    assert [b.size for b in r["cdr_bits"]] == list(g["bit_counts"])
    assert np.array_equal(np.concatenate(r["cdr_bits"]), g["cdr_bits"])
    assert np.array_equal(np.concatenate(r["diff_bits"]), g["diff_bits"])
    assert r["offsets"] == str(g["offsets"])


def test_bit_layer_matches_model():
    R = orclib.RDS()
    g = np.load(os.path.join(GOLD, "rds_bits.npz"))
    n_cdr = 0
    for k in range(6):
        for bc in (0, 3):
            got = R.cdr(g[f"cdr_in_{k}"], 26, bc)
            assert np.array_equal(got, g[f"cdr_out_{k}_bc{bc}"]), (k, bc)
            n_cdr += 1
    assert n_cdr == 12
    assert np.array_equal(R.diff_decode(g["diff_in"]), g["diff_out"])
    for k in range(6):
        off, idx = R.framesync(g[f"fs_in_{k}"])
        assert [off, str(idx)] == list(g[f"fs_out_{k}"]), k
    for name in "ABCcD":
        assert np.array_equal(R.syndrome(g[f"synd_in_{name}"]), g[f"synd_out_{name}"])


def test_cdr_with_carried_state_matches_model():
    """The model's CDR with its to_pass_on_state kept from block to block (fmSupportLib.py:104-106,
    :178-189), on hand-made inputs (odd leftover -> pairing branch :117-125) and on the RRC output
    of the golden capture."""
    R = orclib.RDS()
    g = np.load(os.path.join(GOLD, "rds_bits.npz"))
    for k in range(6):
        for j, (p0, start, prev) in enumerate([(0.7, 5, 3), (-0.4, 20, 4), (0.0, 0, 1)]):
            st = np.array([p0, 0.0, start, prev], np.float64)
            got = R.cdr_state(g[f"cdr_in_{k}"], 26, 2, st)
            assert np.array_equal(got, g[f"cdrs_out_{k}_{j}"]), (k, j)
            assert np.array_equal(st, g[f"cdrs_state_{k}_{j}"]), (k, j, st, g[f"cdrs_state_{k}_{j}"])
    for mode, n_ref, block_if, n_blocks in ((0, 15, 9600, 8), (2, 12, 19200, 2)):
        gm = np.load(os.path.join(GOLD, f"rds_mode{mode}.npz"))
        fm = fm_demod_of(mode, n_ref)[:block_if * n_blocks]
        r = R.run_chain(fm, mode, block_if, keep=(), cdr_carry=True)
        assert [b.size for b in r["cdr_bits"]] == list(gm["carry_counts"])
        assert np.array_equal(np.concatenate(r["cdr_bits"]), gm["carry_bits"])


def test_offset_word_syndromes():
    """Known answers of the RDS standard (the reference's doc/3dy4-project-2022.pdf p.21 and
    fmSupportLib.py:62-91): a block of zero information bits whose check bits are the offset
    word itself has that offset's syndrome."""
    R = orclib.RDS()
    kat = {"A": ("0011111100", "1111011000"), "B": ("0110011000", "1111010100"),
           "C": ("0101101000", "1001011100"), "c": ("1101010000", "1111001100"),
           "D": ("0110110100", "1001011000")}
    for name, (word, synd) in kat.items():
        blk = np.array([0] * 16 + [int(c) for c in word], np.uint8)
        assert "".join(str(int(b)) for b in R.syndrome(blk)) == synd, name
        off, _ = R.framesync(np.concatenate((blk, np.zeros(30, np.uint8))))
        assert off == name
