"""ctypes access to oracle/aux_oracle.c (test infrastructure only)."""
import ctypes as C

import numpy as np

import orclib

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")


def _lib():
    L = orclib.ORC().lib
    L.orc_psd.restype = C.c_int
    L.orc_psd.argtypes = [f32p, C.c_size_t, C.c_float, f32p, f32p]
    L.orc_deemphasis.restype = None
    L.orc_deemphasis.argtypes = [i16p, C.c_size_t, C.c_int, C.c_float, C.c_float, f32p]
    L.orc_channelize.restype = None
    L.orc_channelize.argtypes = [u8p, C.c_size_t, C.c_int, f32p, C.c_int, C.c_float, u8p, C.c_size_t]
    return L


def psd(x, fs):
    x = np.ascontiguousarray(x, np.float32)
    freq, out = np.zeros(256, np.float32), np.zeros(256, np.float32)
    _lib().orc_psd(x, x.size, fs, freq, out)
    return freq, out


def ref_psd(x, fs):
    ref = orclib.REF()
    ref.lib.ref_psd.restype = C.c_int
    ref.lib.ref_psd.argtypes = [f32p, C.c_size_t, C.c_float, f32p, f32p]
    x = np.ascontiguousarray(x, np.float32)
    freq, out = np.zeros(256, np.float32), np.zeros(256, np.float32)
    ref.lib.ref_psd(x, x.size, fs, freq, out)
    return freq, out


def deemphasis(pcm_row, channels, fs, tau, state):
    """pcm_row: int16 1-D (frames * channels), filtered in place; state float32[channels] carried."""
    _lib().orc_deemphasis(pcm_row, pcm_row.size // channels, channels, fs, tau, state)
    return pcm_row


def channelize(iq, M, h, gain):
    iq = np.ascontiguousarray(iq, np.uint8)
    n_in = iq.size // 2
    n_out = n_in // M
    out = np.zeros((M, 2 * n_out), np.uint8)
    _lib().orc_channelize(iq, n_in, M, np.ascontiguousarray(h, np.float32), h.size, gain, out, n_out)
    return out
