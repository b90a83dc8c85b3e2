"""Generate tests/golden/*.npz from the UNMODIFIED reference compiled into
oracle/_ref/libfmref.so (oracle/Makefile).  Run in the build container, where
/root/reference exists:   python tests/golden/make_golden.py

chain_mode{m}.npz   2 reference blocks of synthetic stereo I/Q for mode m, and for
                    each (channels, tap set) the reference's PCM plus the SHA-256 of every
                    float intermediate (raw little-endian float32 bytes).
libm.npz            arguments and results of the host libm's atan2f / sincosf / cosf
                    (glibc 2.39, x86-64, FMA variant) -- what fmPLL calls.
design.npz          the filter coefficient sets the receiver designs.
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
from sdr_b200 import siggen  # noqa: E402

TAPSETS = {"F": (151, 101, 151), "S": (13, 13, 13)}


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    R = orclib.REF()
    assert R is not None, "needs oracle/_ref/libfmref.so (build container only)"
    for mode in range(4):
        iq = siggen.make_capture(100 + mode, mode, 2, "stereo")
        out = {"iq": iq}
        for ch in (1, 2):
            for tname, taps in TAPSETS.items():
                pcm, t = R.run_chain(iq, mode, ch, *taps)
                key = f"c{ch}_{tname}"
                out[key + "_pcm"] = pcm
                for name, arr in t.items():
                    out[f"{key}_sha_{name}"] = np.array(sha(arr))
                    out[f"{key}_n_{name}"] = np.array(arr.size)
        np.savez_compressed(os.path.join(HERE, f"chain_mode{mode}.npz"), **out)
    rng = np.random.default_rng(7)
    n = 4096
    x = np.concatenate([rng.uniform(-4, 4, n), rng.uniform(-130, 130, n),
                        rng.uniform(-4e5, 4e5, n), rng.standard_normal(n) * 1e-5]).astype(np.float32)
    s = np.empty_like(x); c = np.empty_like(x); c2 = np.empty_like(x)
    R.lib.ref_libm_sincosf(x, x.size, s, c)
    R.lib.ref_libm_cosf(x, x.size, c2)
    ay = (rng.standard_normal(4 * n) * 10.0 ** rng.uniform(-6, 2, 4 * n)).astype(np.float32)
    ax = (rng.standard_normal(4 * n) * 10.0 ** rng.uniform(-6, 2, 4 * n)).astype(np.float32)
    ay[:8] = [0, -0.0, 1, -1, 0, 0, 3, -3]
    ax[:8] = [1, 1, 0, 0, -1, -0.0, 1, 1]
    at = np.empty_like(ay)
    R.lib.ref_libm_atan2f(ay, ax, ay.size, at)
    np.savez_compressed(os.path.join(HERE, "libm.npz"), x=x, sin=s, cos=c, cosf=c2, ay=ay, ax=ax, atan2=at)
    d = {}
    for name, args in {"rf_m0": (2.4e6, 1e5, 151), "rf_m1": (1.44e6, 1e5, 151), "rf_m3": (9.6e5, 1e5, 151),
                       "rf_s": (2.4e6, 1e5, 13), "audio_m0": (240000, 16000, 101),
                       "audio_m1": (288000, 16000, 101), "audio_m2": (240000 * 147, 16000, 101 * 147),
                       "audio_m3": (320000 * 441, 16000, 101 * 441)}.items():
        d["lpf_" + name] = R.lpf(*args)
        d["lpf_" + name + "_args"] = np.array(args, np.float64)
    for name, args in {"pilot_240": (240000, 18.5e3, 19.5e3, 151), "stereo_240": (240000, 22e3, 54e3, 151),
                       "pilot_288": (288000, 18.5e3, 19.5e3, 151), "stereo_320_s": (320000, 22e3, 54e3, 13)}.items():
        d["bpf_" + name] = R.bpf(*args)
        d["bpf_" + name + "_args"] = np.array(args, np.float64)
    np.savez_compressed(os.path.join(HERE, "design.npz"), **d)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
