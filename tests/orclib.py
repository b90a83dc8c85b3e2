"""ctypes access to the two CPU checkers (test infrastructure only).

``ORC``  -- oracle/libfm_oracle.so, the plain-C restatement (always available).
``REF``  -- oracle/_ref/libfmref.so, the reference's own sources compiled unmodified
            (available in the build container and, prebuilt, on the GPU box).
Both expose the same calls with an ``orc_`` / ``ref_`` prefix.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

TAP_NAMES = ["i_filt", "q_filt", "demod", "allpass", "stereo_filt", "carrier_filt", "nco",
             "mixer", "audio_filt", "stereo_final"]

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")


class Config(C.Structure):
    _fields_ = [("mode", C.c_int), ("channels", C.c_int), ("rf_taps", C.c_int),
                ("audio_taps", C.c_int), ("stereo_taps", C.c_int)]


def build_oracle() -> None:
    """Compile the checkers if they are missing (make is a no-op when up to date)."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True, capture_output=True)


class CpuLib:
    def __init__(self, path: str, prefix: str):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        L, p = self.lib, prefix

        def fn(name, res, args):
            f = getattr(L, p + name)
            f.restype, f.argtypes = res, args
            return f

        self.lpf_design = fn("lpf_design", None, [C.c_float, C.c_float, C.c_ushort, f32p])
        self.bpf_design = fn("bpf_design", None, [C.c_float, C.c_float, C.c_float, C.c_ushort, f32p])
        self.u8_to_f32 = fn("u8_to_f32", None, [u8p, C.c_size_t, f32p])
        self.fir_block = fn("fir_block", None, [f32p, f32p, C.c_size_t, f32p, C.c_size_t, f32p])
        self.fir_decim = fn("fir_decim", None, [f32p, f32p, C.c_size_t, f32p, C.c_size_t, f32p, C.c_uint])
        self.fir_resample = fn("fir_resample", None,
                               [f32p, f32p, C.c_size_t, f32p, C.c_size_t, f32p, C.c_uint, C.c_uint])
        self.fm_demod = fn("fm_demod", None, [f32p, f32p, f32p, C.c_size_t,
                                              C.POINTER(C.c_float), C.POINTER(C.c_float)])
        self.allpass = fn("allpass", None, [f32p, C.c_size_t, f32p, C.c_size_t, f32p])
        self.pll = fn("pll", None, [f32p, C.c_size_t, f32p, f32p, C.c_float, C.c_float,
                                    C.c_float, C.c_float, C.c_float])
        self.pcm16 = fn("pcm16", C.c_int16, [C.c_float])
        self.chain_create = fn("chain_create", C.c_void_p, [C.POINTER(Config)])
        self.chain_destroy = fn("chain_destroy", None, [C.c_void_p])
        self.chain_reset = fn("chain_reset", None, [C.c_void_p])
        self.chain_process = fn("chain_process", C.c_size_t,
                                [C.c_void_p, u8p, C.c_size_t, i16p, C.c_int])
        self.chain_tap = fn("chain_tap", C.POINTER(C.c_float),
                            [C.c_void_p, C.c_int, C.POINTER(C.c_size_t)])
        self.chain_clear_taps = fn("chain_clear_taps", None, [C.c_void_p])

    # -- convenience -------------------------------------------------------
    def lpf(self, Fs, Fc, n):
        h = np.zeros(n, np.float32)
        self.lpf_design(Fs, Fc, n, h)
        return h

    def bpf(self, Fs, Fb, Fe, n):
        h = np.zeros(n, np.float32)
        self.bpf_design(Fs, Fb, Fe, n, h)
        return h

    def run_chain(self, iq: np.ndarray, mode: int, channels: int, rf_taps=151, audio_taps=101,
                  stereo_taps=151, keep_taps=True):
        """Process one capture; returns (pcm int16 array, {tap name: float32 array})."""
        from_cfg = Config(mode, channels, rf_taps, audio_taps, stereo_taps)
        h = self.chain_create(C.byref(from_cfg))
        if not h:
            raise ValueError("bad chain config")
        try:
            iq = np.ascontiguousarray(iq, np.uint8)
            pcm = np.zeros(iq.size // 8 + 16, np.int16)  # generous: >= 2 * n_audio
            n = self.chain_process(h, iq, iq.size, pcm, 1 if keep_taps else 0)
            taps = {}
            if keep_taps:
                for sid, name in enumerate(TAP_NAMES):
                    cnt = C.c_size_t(0)
                    ptr = self.chain_tap(h, sid, C.byref(cnt))
                    if cnt.value:
                        taps[name] = np.ctypeslib.as_array(ptr, shape=(cnt.value,)).copy()
            return pcm[:n].copy(), taps
        finally:
            self.chain_destroy(h)


class RefLib(CpuLib):
    def __init__(self, path):
        super().__init__(path, "ref_")
        L = self.lib
        L.ref_libm_atan2f.restype = None
        L.ref_libm_atan2f.argtypes = [f32p, f32p, C.c_size_t, f32p]
        L.ref_libm_sincosf.restype = None
        L.ref_libm_sincosf.argtypes = [f32p, C.c_size_t, f32p, f32p]
        L.ref_libm_cosf.restype = None
        L.ref_libm_cosf.argtypes = [f32p, C.c_size_t, f32p]


class OrcLib(CpuLib):
    def __init__(self, path):
        super().__init__(path, "orc_")
        L = self.lib
        L.orc_atan2f.restype = C.c_float
        L.orc_atan2f.argtypes = [C.c_float, C.c_float]
        L.orc_cosf.restype = C.c_float
        L.orc_cosf.argtypes = [C.c_float]
        L.orc_sincosf.restype = None
        L.orc_sincosf.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orc_atan2f_batch.restype = None
        L.orc_atan2f_batch.argtypes = [f32p, f32p, C.c_size_t, f32p]
        L.orc_sincosf_batch.restype = None
        L.orc_sincosf_batch.argtypes = [f32p, C.c_size_t, f32p, f32p]


_orc = None
_ref = None


def ORC() -> OrcLib:
    global _orc
    if _orc is None:
        path = os.path.join(ORACLE_DIR, "libfm_oracle.so")
        if not os.path.exists(path):
            build_oracle()
        _orc = OrcLib(path)
    return _orc


def REF() -> RefLib | None:
    """The compiled reference, or None when it is not available on this host."""
    global _ref
    if _ref is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libfmref.so")
        if not os.path.exists(path) and os.path.isdir("/root/reference/src"):
            build_oracle()
        if not os.path.exists(path):
            return None
        _ref = RefLib(path)
    return _ref
