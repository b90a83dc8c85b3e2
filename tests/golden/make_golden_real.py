"""Generate tests/golden/real_iq.npz: the only RECORDED I/Q data the reference ships -- two blocks of
102 400 samples that its front end printed as floats into data/data/pipeData.txt (lines 5 and 7,
values (byte-128)/128 with six significant digits, so the bytes are recovered exactly) -- together
with what the UNMODIFIED reference (oracle/_ref/libfmref.so) makes of them in mode 0: PCM for mono
and stereo with both tap sets, and the SHA-256 of every float intermediate.
Run in the build container:   python tests/golden/make_golden_real.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orclib  # noqa: E402
import sdr_b200  # noqa: E402,F401

SRC = "/root/reference/data/data/pipeData.txt"
TAPSETS = {"F": (151, 101, 151), "S": (13, 13, 13)}


def main():
    lines = open(SRC).read().splitlines()
    blocks = []
    for ln in (4, 6):  # 0-based: lines 5 and 7
        v = np.array(lines[ln].split(), dtype=np.float64)
        assert v.size == 102400
        b = np.rint(v * 128.0) + 128.0
        assert np.all(np.abs(v - (b - 128.0) / 128.0) < 6e-6) and b.min() >= 0 and b.max() <= 255
        blocks.append(b.astype(np.uint8))
    iq = np.concatenate(blocks)
    R = orclib.REF()
    assert R is not None, "needs oracle/_ref/libfmref.so (build container only)"
    out = {"iq": iq}
    for ch in (1, 2):
        for tname, taps in TAPSETS.items():
            pcm, t = R.run_chain(iq, 0, ch, *taps)
            key = f"c{ch}_{tname}"
            out[key + "_pcm"] = pcm
            for name, arr in t.items():
                out[f"{key}_sha_{name}"] = np.array(hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest())
    np.savez_compressed(os.path.join(HERE, "real_iq.npz"), **out)
    print("bytes:", iq.size, "range", iq.min(), iq.max(), "PCM rms (mono, F):",
          float(np.sqrt(np.mean(out["c1_F_pcm"].astype(np.float64) ** 2))))


if __name__ == "__main__":
    main()
