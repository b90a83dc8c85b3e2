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


# ---------------------------------------------------------------------------
# RDS oracle (oracle/rds_oracle.c: restatement of model/fmRDS.py + fmSupportLib.py)
# ---------------------------------------------------------------------------
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")

RDS_TAP_NAMES = ["channel_filt", "carrier_filt", "pll_i", "pll_q", "mixer_i", "mixer_q",
                 "resampler_i", "resampler_q", "rrc_i", "rrc_q"]
RDS_PARAMS = {0: dict(U=247, D=960, sps=26, block_if=9600),
              2: dict(U=817, D=1920, sps=43, block_if=1536000)}


class RdsOracle:
    """ctypes view of the rdo_* functions; lives in the same library as the FM oracle."""

    def __init__(self):
        L = ORC().lib
        self.lib = L
        L.rdo_bandpass.restype = None
        L.rdo_bandpass.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, f64p]
        L.rdo_lowpass.restype = None
        L.rdo_lowpass.argtypes = [C.c_int, C.c_double, C.c_double, f64p]
        L.rdo_rrc.restype = None
        L.rdo_rrc.argtypes = [C.c_double, C.c_int, f64p]
        L.rdo_fir.restype = None
        L.rdo_fir.argtypes = [f64p, C.c_size_t, f64p, C.c_int, f64p, f64p]
        L.rdo_allpass.restype = None
        L.rdo_allpass.argtypes = [f64p, C.c_size_t, f64p, C.c_int, f64p]
        L.rdo_pll.restype = None
        L.rdo_pll.argtypes = [f64p, C.c_size_t, C.c_double, C.c_double, f64p, C.c_double,
                              C.c_double, C.c_double, f64p, f64p]
        L.rdo_resample.restype = None
        L.rdo_resample.argtypes = [f64p, C.c_size_t, f64p, C.c_int, f64p, C.c_int, C.c_int, f64p]
        L.rdo_cdr.restype = C.c_int
        L.rdo_cdr.argtypes = [f64p, C.c_int, C.c_int, C.c_int, u8p, C.c_int]
        L.rdo_cdr_state.restype = C.c_int
        L.rdo_cdr_state.argtypes = [f64p, C.c_int, C.c_int, C.c_int, f64p, u8p, C.c_int]
        L.rdo_chain_set_cdr_carry.restype = None
        L.rdo_chain_set_cdr_carry.argtypes = [C.c_void_p, C.c_int]
        L.rdo_diff_decode.restype = None
        L.rdo_diff_decode.argtypes = [u8p, C.c_int, u8p]
        L.rdo_syndrome.restype = None
        L.rdo_syndrome.argtypes = [u8p, u8p]
        L.rdo_framesync.restype = C.c_char
        L.rdo_framesync.argtypes = [u8p, C.c_int, C.POINTER(C.c_int)]
        L.rdo_chain_create.restype = C.c_void_p
        L.rdo_chain_create.argtypes = [C.c_int, C.c_int]
        L.rdo_chain_destroy.restype = None
        L.rdo_chain_destroy.argtypes = [C.c_void_p]
        L.rdo_chain_block.restype = None
        L.rdo_chain_block.argtypes = [C.c_void_p, f64p]
        L.rdo_chain_tap.restype = C.POINTER(C.c_double)
        L.rdo_chain_tap.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t)]
        L.rdo_chain_bits.restype = C.c_int
        L.rdo_chain_bits.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_uint8)),
                                     C.POINTER(C.POINTER(C.c_uint8))]
        L.rdo_chain_offset.restype = C.c_char
        L.rdo_chain_offset.argtypes = [C.c_void_p]

    # -- design ------------------------------------------------------------
    def bandpass(self, ntaps, Fs, Fb, Fe):
        h = np.empty(ntaps, np.float64)
        self.lib.rdo_bandpass(ntaps, Fs, Fb, Fe, h)
        return h

    def lowpass(self, ntaps, Fs, Fc):
        h = np.empty(ntaps, np.float64)
        self.lib.rdo_lowpass(ntaps, Fs, Fc, h)
        return h

    def rrc(self, Fs, ntaps):
        h = np.empty(ntaps, np.float64)
        self.lib.rdo_rrc(Fs, ntaps, h)
        return h

    # -- bit layer -----------------------------------------------------------
    def cdr(self, x, sps, block_count):
        x = np.ascontiguousarray(x, np.float64)
        bits = np.zeros(x.size // sps + 4, np.uint8)
        n = self.lib.rdo_cdr(x, x.size, sps, block_count, bits, bits.size)
        return bits[:n].copy()

    def cdr_state(self, x, sps, block_count, state):
        """state: float64[4] = pair[0], pair[1], start, prev_size; updated in place."""
        x = np.ascontiguousarray(x, np.float64)
        bits = np.zeros(x.size // sps + 4, np.uint8)
        n = self.lib.rdo_cdr_state(x, x.size, sps, block_count, state, bits, bits.size)
        return bits[:n].copy()

    def diff_decode(self, bits):
        bits = np.ascontiguousarray(bits, np.uint8)
        out = np.zeros_like(bits)
        self.lib.rdo_diff_decode(bits, bits.size, out)
        return out

    def syndrome(self, d26):
        d = np.ascontiguousarray(d26, np.uint8)
        s = np.zeros(10, np.uint8)
        self.lib.rdo_syndrome(d, s)
        return s

    def framesync(self, bits):
        bits = np.ascontiguousarray(bits, np.uint8)
        idx = C.c_int(0)
        o = self.lib.rdo_framesync(bits if bits.size else np.zeros(1, np.uint8), bits.size,
                                   C.byref(idx))
        return o.decode(), idx.value

    # -- chain -----------------------------------------------------------------
    def run_chain(self, fm_demod, mode, block_if=None, keep=("rrc_i", "rrc_q"), cdr_carry=False):
        """fm_demod: float array, a whole number of blocks.  Returns a dict with the kept
        stages concatenated over blocks, 'cdr_bits' / 'diff_bits' (lists per block) and
        'offsets' (one character per block)."""
        block_if = block_if or RDS_PARAMS[mode]["block_if"]
        x = np.ascontiguousarray(fm_demod, np.float64)
        assert x.size % block_if == 0
        h = self.lib.rdo_chain_create(mode, block_if)
        if not h:
            raise ValueError("RDS is defined for modes 0 and 2 only")
        self.lib.rdo_chain_set_cdr_carry(h, 1 if cdr_carry else 0)
        out = {k: [] for k in keep}
        out.update(cdr_bits=[], diff_bits=[], offsets="")
        try:
            for b in range(x.size // block_if):
                self.lib.rdo_chain_block(h, x[b * block_if:(b + 1) * block_if])
                for k in keep:
                    cnt = C.c_size_t(0)
                    ptr = self.lib.rdo_chain_tap(h, RDS_TAP_NAMES.index(k), C.byref(cnt))
                    out[k].append(np.ctypeslib.as_array(ptr, shape=(cnt.value,)).copy())
                p1 = C.POINTER(C.c_uint8)()
                p2 = C.POINTER(C.c_uint8)()
                n = self.lib.rdo_chain_bits(h, C.byref(p1), C.byref(p2))
                out["cdr_bits"].append(np.ctypeslib.as_array(p1, shape=(max(n, 1),))[:n].copy())
                out["diff_bits"].append(np.ctypeslib.as_array(p2, shape=(max(n, 1),))[:n].copy())
                out["offsets"] += self.lib.rdo_chain_offset(h).decode()
        finally:
            self.lib.rdo_chain_destroy(h)
        for k in keep:
            out[k] = np.concatenate(out[k])
        return out


_rds = None


def RDS() -> RdsOracle:
    global _rds
    if _rds is None:
        _rds = RdsOracle()
    return _rds
