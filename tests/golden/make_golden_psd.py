"""Golden vectors for the PSD diagnostics op: outputs of the reference's own estimatePSD
(src/fourier.cpp:44-126, compiled unmodified into oracle/_ref/libfmref.so) on seeded inputs.
Run in the build container (needs /root/reference): python tests/golden/make_golden_psd.py"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orclib  # noqa: E402

ref = orclib.REF()
assert ref is not None, "needs oracle/_ref/libfmref.so (built from /root/reference)"
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
ref.lib.ref_psd.restype = C.c_int
ref.lib.ref_psd.argtypes = [f32p, C.c_size_t, C.c_float, f32p, f32p]
rng = np.random.default_rng(0x5D2)
cases = {}
for name, (fs, n) in {"audio48k": (48000.0, 512 * 9 + 100), "if240k": (240000.0, 512 * 10), "one_segment": (44100.0, 600)}.items():
    t = np.arange(n) / fs
    x = (0.3 * np.sin(2 * np.pi * 0.02 * fs * t) + 0.1 * np.sin(2 * np.pi * 0.31 * fs * t + 1.0) +
         0.02 * rng.standard_normal(n)).astype(np.float32)
    freq, psd = np.zeros(256, np.float32), np.zeros(256, np.float32)
    ref.lib.ref_psd(x, x.size, fs, freq, psd)
    cases[f"{name}_x"], cases[f"{name}_fs"], cases[f"{name}_freq"], cases[f"{name}_psd"] = x, np.float32(fs), freq, psd
np.savez_compressed(os.path.join(HERE, "psd.npz"), **cases)
print("wrote psd.npz:", sorted(cases))
