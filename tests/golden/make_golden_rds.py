"""Generate tests/golden/rds_*.npz from the UNMODIFIED reference model: this script imports
/root/reference/model/fmSupportLib.py (numpy only) and scipy.signal.lfilter, exactly the
functions model/fmRDS.py:222-276 calls, and runs that block loop headless (fmRDS.py itself
needs matplotlib and a recorded capture, both absent).  Run in the build container:

    python tests/golden/make_golden_rds.py

Input: the synthetic capture siggen.make_capture(200, mode, ..., "rds"); its fm_demod comes
from the reference's own C++ front end (oracle/_ref), cast to float64.  The test regenerates
the same fm_demod with the (bit-identical, separately pinned) C oracle, so it is not stored.

rds_design.npz   coefficient sets (bandPass, impResponse, RRC) for modes 0 and 2
rds_mode0.npz    8 blocks of 9600 IF samples (fmRDS.py:149): RRC I/Q output in full, every
                 16th sample of the other stages for the last block, CDR / differential bits
                 and the frame synchroniser's result per block
rds_mode2.npz    2 blocks of 19200 IF samples (the model's own mode-2 block is 1 536 000 IF
                 samples, too slow for its pure-Python resampler; the chain is a streaming
                 one, only the CDR window depends on the block length)
rds_bits.npz     CDR / diff_decoding / framesync on hand-made inputs that reach the branches
                 the synthetic capture does not (inversions, restarts, all five offset words)
"""
import math
import os
import sys

import numpy as np
from scipy import signal

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference/model")
import fmSupportLib as M  # noqa: E402  the reference model
import orclib  # noqa: E402
import sdr_b200  # noqa: E402,F401
from sdr_b200 import siggen  # noqa: E402

PARAMS = {0: dict(U=247, D=960, sps=26), 2: dict(U=817, D=1920, sps=43)}  # fmRDS.py:55-75
IF_FS = 240000


def coeffs(mode):
    p = PARAMS[mode]
    return dict(  # fmRDS.py:122-125
        chan=M.bandPass(151, IF_FS, 54e3, 60e3),
        carr=M.bandPass(151, IF_FS, 113.5e3, 114.5e3),
        rs=M.impResponse(101 * p["U"], IF_FS * p["U"], 3e3),
        rrc=M.impulseResponseRootRaisedCosine(2375 * p["sps"], 101))


def model_chain(fm_demod, mode, block_if):
    """fmRDS.py:158-276 from fm_demod onwards, one block at a time."""
    p = PARAMS[mode]
    c = coeffs(mode)
    st_pll = [0.0, 0.0, 1.0, 0.0, 1.0, 0, 1.0]
    st_chan = np.zeros(150); st_carr = np.zeros(150); st_ap = np.zeros(75)
    st_rs = np.zeros(101 * p["U"] - 1); st_rs2 = np.zeros(101 * p["U"] - 1)
    st_rrc = np.zeros(100); st_rrc2 = np.zeros(100)
    decoded = np.array([])
    out = dict(rrc_i=[], rrc_q=[], cdr=[], diff=[], offsets="", last={}, sps=p["sps"])
    for bc in range(fm_demod.size // block_if):
        x = fm_demod[bc * block_if:(bc + 1) * block_if]
        chan, st_chan = signal.lfilter(c["chan"], 1.0, x, zi=st_chan)
        ap, st_ap = M.allPass(chan, st_ap)
        sq = chan * chan
        carr, st_carr = signal.lfilter(c["carr"], 1.0, sq, zi=st_carr)
        pll_i, pll_q, st_pll = M.fmPll(carr, 114e3, IF_FS, st_pll, ncoScale=0.5,
                                        phaseAdjust=(3 * math.pi / 8), normBandwidth=0.002)
        mix = pll_i[:-1] * ap * 2
        rs, st_rs = M.convolveBlockResampleFIR(mix, c["rs"], st_rs, p["D"], p["U"])
        rrc, st_rrc = signal.lfilter(c["rrc"], 1.0, rs, zi=st_rrc)
        mix2 = pll_q[:-1] * ap * 2
        rs2, st_rs2 = M.convolveBlockResampleFIR(mix2, c["rs"], st_rs2, p["D"], p["U"])
        rrc2, st_rrc2 = signal.lfilter(c["rrc"], 1.0, rs2, zi=st_rrc2)
        state = [np.zeros(2), 158, 0]  # fmRDS.py:257-260
        bits, state = M.CDR(rrc, p["sps"], state, bc)
        diff = M.diff_decoding(bits)
        decoded = np.concatenate((decoded, diff))
        off, idx = M.framesync(decoded)
        decoded = decoded[idx:]
        out["rrc_i"].append(rrc); out["rrc_q"].append(rrc2)
        out["cdr"].append(bits.astype(np.uint8)); out["diff"].append(diff.astype(np.uint8))
        out["offsets"] += {"C_apos": "c"}.get(off, off)
        out["last"] = dict(channel_filt=chan, carrier_filt=carr, pll_i=pll_i, pll_q=pll_q,
                           mixer_i=mix, mixer_q=mix2, resampler_i=rs, resampler_q=rs2)
        print(f"mode {mode} block {bc}: {bits.size} bits, offset {off!r}", flush=True)
    return out


def demod_of(mode, n_ref_blocks):
    R = orclib.REF()
    assert R is not None, "needs oracle/_ref/libfmref.so (build container only)"
    iq = siggen.make_capture(200, mode, n_ref_blocks, "rds")
    _, taps = R.run_chain(iq, mode, 1)
    return taps["demod"].astype(np.float64)


def carried_cdr(rrc_blocks, sps):
    """The model's CDR with its to_pass_on_state kept from block to block (what the function's
    state arguments are for; fmRDS.py:257-260 re-creates the state instead)."""
    state = [np.zeros(2), 158, 0]
    bits_all, counts = [], []
    for bc, rrc in enumerate(rrc_blocks):
        bits, state = M.CDR(rrc, sps, state, bc)
        bits_all.append(bits.astype(np.uint8))
        counts.append(bits.size)
    return dict(carry_bits=np.concatenate(bits_all), carry_counts=np.array(counts),
                carry_state=np.array([state[0][0], state[0][1], state[1], state[2]], dtype=np.float64))


def pack(res):
    d = dict(rrc_i=np.concatenate(res["rrc_i"]), rrc_q=np.concatenate(res["rrc_q"]),
             offsets=np.array(res["offsets"]),
             bit_counts=np.array([b.size for b in res["cdr"]]),
             cdr_bits=np.concatenate(res["cdr"]), diff_bits=np.concatenate(res["diff"]))
    for k, v in res["last"].items():
        d["last16_" + k] = v[::16].copy()
    d.update(carried_cdr(res["rrc_i"], res["sps"]))
    return d


def bit_layer_cases():
    rng = np.random.default_rng(11)
    out = {}
    # CDR on hand-made symbol streams (26 samples per symbol, like mode 0)
    sps, n = 26, 2470
    cases = []
    for k in range(6):
        x = rng.standard_normal(n) * 0.05
        chips = rng.integers(0, 2, 200) * 2 - 1
        if k >= 2:  # biphase pairs -> mostly regular, with a few irregular pairs
            b = rng.integers(0, 2, 100) * 2 - 1
            chips = np.repeat(b, 2) * np.tile([1, -1], 100)
            bad = rng.integers(0, 200, 3 + k)
            chips[bad] *= -1
        amp = [1.0, 0.2, 1.0, 0.25, 0.6, 1.0][k]
        pos = 158 + (26 if k == 5 else 0) + sps * np.arange(89 - (1 if k == 5 else 0))
        x[pos] = amp * chips[:pos.size] * rng.uniform(0.5, 1.5, pos.size)
        cases.append(x)
    for k, x in enumerate(cases):
        for bc in (0, 3):
            st = [np.zeros(2), 158, 0]
            bits, _ = M.CDR(x, sps, st, bc)
            out[f"cdr_in_{k}"] = x
            out[f"cdr_out_{k}_bc{bc}"] = bits.astype(np.uint8)
    # the same inputs with a carried state: an odd number of points left over (pairing branch,
    # fmSupportLib.py:117-125), other start offsets
    for k, x in enumerate(cases):
        for j, (p0, start, prev) in enumerate([(0.7, 5, 3), (-0.4, 20, 4), (0.0, 0, 1)]):
            st = [np.array([p0, 0.0]), start, prev]
            bits, st = M.CDR(x, sps, st, 2)
            out[f"cdrs_out_{k}_{j}"] = bits.astype(np.uint8)
            out[f"cdrs_state_{k}_{j}"] = np.array([st[0][0], st[0][1], st[1], st[2]], dtype=np.float64)
    # differential decoding
    mb = rng.integers(0, 2, 64).astype(np.float64)
    out["diff_in"] = mb.astype(np.uint8)
    out["diff_out"] = M.diff_decoding(mb).astype(np.uint8)
    # frame synchroniser: random bits with valid blocks (all five offset words) spliced in.
    # A valid block = 16 information bits + (checkword XOR offset word); with the parity matrix
    # H = [I10; P] the checkword of info bits m is chosen so that [m | c] . H = offset syndrome.
    Hm = np.array([[1,0,0,0,0,0,0,0,0,0],[0,1,0,0,0,0,0,0,0,0],[0,0,1,0,0,0,0,0,0,0],
                   [0,0,0,1,0,0,0,0,0,0],[0,0,0,0,1,0,0,0,0,0],[0,0,0,0,0,1,0,0,0,0],
                   [0,0,0,0,0,0,1,0,0,0],[0,0,0,0,0,0,0,1,0,0],[0,0,0,0,0,0,0,0,1,0],
                   [0,0,0,0,0,0,0,0,0,1],[1,0,1,1,0,1,1,1,0,0],[0,1,0,1,1,0,1,1,1,0],
                   [0,0,1,0,1,1,0,1,1,1],[1,0,1,0,0,0,0,1,1,1],[1,1,1,0,0,1,1,1,1,1],
                   [1,1,0,0,0,1,0,0,1,1],[1,1,0,1,0,1,0,1,0,1],[1,1,0,1,1,1,0,1,1,0],
                   [0,1,1,0,1,1,1,0,1,1],[1,0,0,0,0,0,0,0,0,1],[1,1,1,1,0,1,1,1,0,0],
                   [0,1,1,1,1,0,1,1,1,0],[0,0,1,1,1,1,0,1,1,1],[1,0,1,0,1,0,0,1,1,1],
                   [1,1,1,0,0,0,1,1,1,1],[1,1,0,0,0,1,1,0,1,1]])
    synd = dict(A=[1,1,1,1,0,1,1,0,0,0], B=[1,1,1,1,0,1,0,1,0,0], C=[1,0,0,1,0,1,1,1,0,0],
                c=[1,1,1,1,0,0,1,1,0,0], D=[1,0,0,1,0,1,1,0,0,0])

    def block_with(s):
        # first 10 positions meet the identity rows: choose the last 16 freely, solve the first 10
        tail = rng.integers(0, 2, 16)
        head = (np.array(s) + tail @ Hm[10:]) % 2
        return np.concatenate((head, tail))

    for k, order in enumerate(["A", "ABCD", "AcD", "", "DDDD", "BA"]):
        parts = [rng.integers(0, 2, int(rng.integers(0, 40)))]
        for o in order:
            parts.append(block_with(synd[o]))
            if rng.integers(0, 2):
                parts.append(rng.integers(0, 2, int(rng.integers(0, 30))))
        parts.append(rng.integers(0, 2, int(rng.integers(0, 60))))
        d = np.concatenate(parts).astype(np.float64)
        off, idx = M.framesync(d)
        out[f"fs_in_{k}"] = d.astype(np.uint8)
        out[f"fs_out_{k}"] = np.array([{"C_apos": "c"}.get(off, off), str(idx)])
    # syndromes of single blocks
    for name, s in synd.items():
        blk = block_with(s).astype(np.float64)
        out[f"synd_in_{name}"] = blk.astype(np.uint8)
        out[f"synd_out_{name}"] = M.matrixMult(blk, Hm.tolist()).astype(np.uint8)
    return out


def main():
    d = {}
    for mode in (0, 2):
        for k, v in coeffs(mode).items():
            d[f"m{mode}_{k}"] = v if v.size <= 151 else v[::97].copy()
            d[f"m{mode}_{k}_sum"] = np.array([v.sum(), np.abs(v).sum()])
    np.savez_compressed(os.path.join(HERE, "rds_design.npz"), **d)
    # mode 0: 15 reference blocks of 102 400 B = 8 RDS blocks of 192 000 B
    fm = demod_of(0, 15)
    np.savez_compressed(os.path.join(HERE, "rds_mode0.npz"), **pack(model_chain(fm, 0, 9600)))
    # mode 2: 12 reference blocks of 112 000 B = 67 200 IF samples -> first 2 x 19 200
    fm = demod_of(2, 12)[:38400]
    np.savez_compressed(os.path.join(HERE, "rds_mode2.npz"), **pack(model_chain(fm, 2, 19200)))
    np.savez_compressed(os.path.join(HERE, "rds_bits.npz"), **bit_layer_cases())


if __name__ == "__main__":
    main()
