"""Python binding of the B200-native FM receiver DSP path (ctypes over the C ABI).

The product is the shared library ``libsdr_b200.so`` (CUDA kernels for sm_100a behind
``include/sdr_b200.h``); this module only marshals numpy arrays / raw device pointers into it
for the tests and the benchmark.  There is no CPU fallback: if the library is missing, or no
sm_100 device is present, the calls raise.

The function names mirror the reference's ``include/filter.h`` (file:line cited per function in
``include/sdr_b200.h``).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsdr_b200.so")

TAP_NAMES = ["i_filt", "q_filt", "demod", "allpass", "stereo_filt", "carrier_filt", "nco",
             "mixer", "audio_filt", "stereo_final"]

VARIANT_EXACT = 0
VARIANT_FAST = 1
VARIANT_MIXED = 2
VARIANT_SCALAR_FIR = 0x100   # flag: scalar form of the exact FIR kernels (same bits; for cross-checks)


class SdrError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"sdr_b200 error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("mode", C.c_int), ("channels", C.c_int), ("rf_taps", C.c_int),
                ("audio_taps", C.c_int), ("stereo_taps", C.c_int), ("batch", C.c_int),
                ("device", C.c_int), ("variant", C.c_int), ("max_bytes_per_channel", C.c_uint64)]


class ModeInfo(C.Structure):
    _fields_ = [("rf_Fs", C.c_int), ("if_Fs", C.c_int), ("audio_Fs", C.c_int),
                ("rf_decim", C.c_int), ("audio_decim", C.c_int), ("audio_upsamp", C.c_int),
                ("block_bytes", C.c_int), ("granule_bytes", C.c_int), ("pcm_per_granule", C.c_int)]


class MultiConfig(C.Structure):
    _fields_ = [("cfg", Config), ("n_devices", C.c_int), ("devices", C.POINTER(C.c_int))]


class ChannelizerConfig(C.Structure):
    _fields_ = [("n_channels", C.c_int), ("taps_per_branch", C.c_int), ("n_wide", C.c_int), ("device", C.c_int),
                ("gain", C.c_float)]


class RdsConfig(C.Structure):
    _fields_ = [("block_if", C.c_int), ("max_pending_blocks", C.c_int), ("keep_nco", C.c_int),
                ("cdr_carry", C.c_int), ("pll_form", C.c_int), ("precision", C.c_int)]


class RdsInfo(C.Structure):
    _fields_ = [("upsamp", C.c_int), ("decim", C.c_int), ("samples_per_symbol", C.c_int),
                ("block_if", C.c_int), ("block_out", C.c_int), ("block_bytes", C.c_int),
                ("max_pending_blocks", C.c_int), ("max_bits_per_block", C.c_int)]


RDS_TAP_NAMES = ["channel_filt", "carrier_filt", "pll_i", "pll_q", "mixer_i", "mixer_q",
                 "resampler_i", "resampler_q", "rrc_i", "rrc_q"]
RDS_FILTER_NAMES = ["channel", "carrier", "resampler", "rrc"]

_f32 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_vp = C.c_void_p

# name -> (restype, argtypes): every symbol include/sdr_b200.h declares
ABI = {
    "sdr_version": (C.c_char_p, []),
    "sdr_last_error": (C.c_char_p, []),
    "sdr_device_count": (C.c_int, []),
    "sdr_lpf_design": (C.c_int, [C.c_float, C.c_float, C.c_ushort, _f32]),
    "sdr_bpf_design": (C.c_int, [C.c_float, C.c_float, C.c_float, C.c_ushort, _f32]),
    "sdr_convolve": (C.c_int, [C.c_int, _f32, _f32, C.c_size_t, _f32, C.c_size_t]),
    "sdr_fir_block": (C.c_int, [C.c_int, _f32, _f32, C.c_size_t, _f32, C.c_size_t, _f32]),
    "sdr_fir_decim": (C.c_int, [C.c_int, _f32, _f32, C.c_size_t, _f32, C.c_size_t, _f32, C.c_uint]),
    "sdr_fir_resample": (C.c_int, [C.c_int, _f32, _f32, C.c_size_t, _f32, C.c_size_t, _f32,
                                   C.c_uint, C.c_uint]),
    "sdr_fm_demod": (C.c_int, [C.c_int, _f32, _f32, _f32, C.c_size_t, C.POINTER(C.c_float),
                               C.POINTER(C.c_float)]),
    "sdr_pll": (C.c_int, [C.c_int, _f32, C.c_size_t, _f32, _f32, C.c_float, C.c_float, C.c_float,
                          C.c_float, C.c_float]),
    "sdr_allpass": (C.c_int, [C.c_int, _f32, C.c_size_t, _f32, C.c_size_t, _f32]),
    "sdr_upsample": (C.c_int, [C.c_int, _f32, C.c_size_t, _f32, C.c_int]),
    "sdr_downsample": (C.c_int, [C.c_int, _f32, _f32, C.c_size_t, C.c_ushort]),
    "sdr_mode_lookup": (C.c_int, [C.c_int, C.c_int, C.POINTER(ModeInfo)]),
    "sdr_pipeline_create": (C.c_int, [C.POINTER(Config), C.POINTER(_vp)]),
    "sdr_pipeline_destroy": (C.c_int, [_vp]),
    "sdr_pipeline_reset": (C.c_int, [_vp]),
    "sdr_pipeline_copy_state": (C.c_int, [_vp, C.c_int, C.c_int]),
    "sdr_pipeline_pcm_count": (C.c_int, [_vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "sdr_pipeline_process_device": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, _vp, C.c_size_t, _vp]),
    "sdr_pipeline_process_host": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, _vp, C.c_size_t]),
    "sdr_pipeline_keep_taps": (C.c_int, [_vp, C.c_int]),
    "sdr_pipeline_tap": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "sdr_pipeline_launch_count": (C.c_int, [_vp, C.POINTER(C.c_uint64), C.c_int]),
    "sdr_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(_vp)]),
    "sdr_host_free": (C.c_int, [_vp]),
    "sdr_pipeline_profile": (C.c_int, [_vp, C.c_int]),
    "sdr_pipeline_kernel_times": (C.c_int, [_vp, C.c_int, C.c_char_p, C.c_size_t,
                                            C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int]),
    "sdr_multi_create": (C.c_int, [C.POINTER(MultiConfig), C.POINTER(_vp)]),
    "sdr_multi_destroy": (C.c_int, [_vp]),
    "sdr_multi_reset": (C.c_int, [_vp]),
    "sdr_multi_layout": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]),
    "sdr_multi_pcm_count": (C.c_int, [_vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "sdr_multi_launch_count": (C.c_int, [_vp, C.POINTER(C.c_uint64), C.c_int]),
    "sdr_multi_process_host": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, _vp, C.c_size_t]),
    "sdr_psd": (C.c_int, [C.c_int, _vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_float, _vp, _vp]),
    "sdr_pipeline_psd": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "sdr_deemph_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.POINTER(_vp)]),
    "sdr_deemph_destroy": (C.c_int, [_vp]),
    "sdr_deemph_reset": (C.c_int, [_vp]),
    "sdr_deemph_process_device": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, _vp]),
    "sdr_deemph_process_host": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t]),
    "sdr_wav_header": (C.c_int, [_vp, C.c_int, C.c_int, C.c_uint64]),
    "sdr_channelizer_create": (C.c_int, [C.POINTER(ChannelizerConfig), C.POINTER(_vp)]),
    "sdr_channelizer_destroy": (C.c_int, [_vp]),
    "sdr_channelizer_reset": (C.c_int, [_vp]),
    "sdr_channelizer_prototype": (C.c_int, [_vp, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "sdr_channelizer_process_device": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, _vp, C.c_size_t, _vp]),
    "sdr_channelizer_process_host": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, _vp, C.c_size_t]),
    "sdr_rds_design": (C.c_int, [C.c_int, C.c_int, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "sdr_rds_create": (C.c_int, [_vp, C.POINTER(RdsConfig), C.POINTER(_vp)]),
    "sdr_rds_destroy": (C.c_int, [_vp]),
    "sdr_rds_info": (C.c_int, [_vp, C.POINTER(RdsInfo)]),
    "sdr_rds_pending": (C.c_int, [_vp, C.POINTER(C.c_size_t)]),
    "sdr_rds_read": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_size_t, C.POINTER(C.c_size_t), _vp, _vp,
                               C.c_size_t, C.POINTER(C.c_size_t)]),
    "sdr_rds_discard": (C.c_int, [_vp]),
    "sdr_rds_tap": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libsdr_b200.so (built by build.py / __graft_entry__.build()); fail loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SdrError(-2, f"{LIB_PATH} is missing: run `python __graft_entry__.py build` "
                               "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in ABI.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise SdrError(rc, lib().sdr_last_error().decode(errors="replace"))


def device_count() -> int:
    return int(lib().sdr_device_count())


def mode_info(mode: int, channels: int = 1) -> ModeInfo:
    mi = ModeInfo()
    _check(lib().sdr_mode_lookup(mode, channels, C.byref(mi)))
    return mi


# ---- filter.h-shaped single operators (numpy in / numpy out) -------------------------------
def impulseResponseLPF(Fs: float, Fc: float, num_taps: int) -> np.ndarray:
    h = np.zeros(num_taps, np.float32)
    _check(lib().sdr_lpf_design(Fs, Fc, num_taps, h))
    return h


def bandPass(Fs: float, Fb: float, Fe: float, num_taps: int) -> np.ndarray:
    h = np.zeros(num_taps, np.float32)
    _check(lib().sdr_bpf_design(Fs, Fb, Fe, num_taps, h))
    return h


def _f(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def convolveFIR(x, h, device: int = 0) -> np.ndarray:
    x, h = _f(x), _f(h)
    y = np.zeros(x.size + h.size - 1, np.float32)
    _check(lib().sdr_convolve(device, y, x, x.size, h, h.size))
    return y


def convolveBlockFIR(x, h, state: np.ndarray, device: int = 0) -> np.ndarray:
    """state (float32, len(h)-1) is updated in place, as in the reference."""
    x, h = _f(x), _f(h)
    y = np.zeros(x.size, np.float32)
    _check(lib().sdr_fir_block(device, y, x, x.size, h, h.size, state))
    return y


def convolveBlockFastFIR(x, h, state: np.ndarray, decim: int, device: int = 0) -> np.ndarray:
    x, h = _f(x), _f(h)
    y = np.zeros(x.size // decim, np.float32)
    _check(lib().sdr_fir_decim(device, y, x, x.size, h, h.size, state, decim))
    return y


def convolveBlockResampleFIR(x, h, state: np.ndarray, decim: int, upsamp: int,
                             device: int = 0) -> np.ndarray:
    x, h = _f(x), _f(h)
    y = np.zeros(x.size * upsamp // decim, np.float32)
    _check(lib().sdr_fir_resample(device, y, x, x.size, h, h.size, state, decim, upsamp))
    return y


def fmDemod(I, Q, prev_i: float, prev_q: float, device: int = 0):
    I, Q = _f(I), _f(Q)
    out = np.zeros(I.size, np.float32)
    pi, pq = C.c_float(prev_i), C.c_float(prev_q)
    _check(lib().sdr_fm_demod(device, out, I, Q, I.size, C.byref(pi), C.byref(pq)))
    return out, pi.value, pq.value


def fmPLL(pll_in, state: np.ndarray, freq, Fs, ncoScale=1.0, phaseAdjust=0.0,
          normBandwidth=0.01, device: int = 0) -> np.ndarray:
    """state (float32[6]) is updated in place; returns ncoOut with len(pll_in)+1 entries."""
    x = _f(pll_in)
    out = np.zeros(x.size + 1, np.float32)
    _check(lib().sdr_pll(device, x, x.size, out, state, freq, Fs, ncoScale, phaseAdjust,
                         normBandwidth))
    return out


def allPass(x, state: np.ndarray, device: int = 0) -> np.ndarray:
    x = _f(x)
    out = np.zeros(x.size, np.float32)
    _check(lib().sdr_allpass(device, x, x.size, state, state.size, out))
    return out


def upsample(x, up_rate: int, device: int = 0) -> np.ndarray:
    x = _f(x)
    out = np.zeros(x.size * up_rate, np.float32)
    _check(lib().sdr_upsample(device, x, x.size, out, up_rate))
    return out


def downsample(x, ds: int, device: int = 0) -> np.ndarray:
    x = _f(x)
    out = np.zeros((x.size + ds - 1) // ds, np.float32)
    _check(lib().sdr_downsample(device, out, x, x.size, ds))
    return out


# ---- batched pipeline ------------------------------------------------------------------------
class Pipeline:
    """project.cpp's receiver for ``batch`` independent captures on one GPU."""

    def __init__(self, mode=0, channels=1, rf_taps=151, audio_taps=101, stereo_taps=151, batch=1,
                 device=0, variant=VARIANT_EXACT, max_bytes_per_channel=0):
        self.cfg = Config(mode, channels, rf_taps, audio_taps, stereo_taps, batch, device, variant,
                          max_bytes_per_channel)
        self.info = mode_info(mode, channels)
        self._h = _vp()
        _check(lib().sdr_pipeline_create(C.byref(self.cfg), C.byref(self._h)))

    def close(self):
        # a follower stage (Rds) holds a pointer to this handle: it goes first
        for f in list(getattr(self, "_followers", [])):
            f.close()
        if self._h:
            lib().sdr_pipeline_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def batch(self) -> int:
        return self.cfg.batch

    def reset(self):
        _check(lib().sdr_pipeline_reset(self._h))

    def keep_taps(self, keep=True):
        _check(lib().sdr_pipeline_keep_taps(self._h, 1 if keep else 0))

    def pcm_count(self, nbytes: int) -> int:
        n = C.c_size_t(0)
        _check(lib().sdr_pipeline_pcm_count(self._h, nbytes, C.byref(n)))
        return n.value

    def launch_count(self, reset=False) -> int:
        n = C.c_uint64(0)
        _check(lib().sdr_pipeline_launch_count(self._h, C.byref(n), 1 if reset else 0))
        return n.value

    def process_host(self, iq: np.ndarray, pcm: np.ndarray | None = None) -> np.ndarray:
        """iq: uint8 [batch, nbytes] (host).  Returns int16 [batch, n_pcm]."""
        iq = np.ascontiguousarray(iq, np.uint8)
        if iq.ndim == 1:
            iq = iq[None, :]
        assert iq.shape[0] == self.batch, "first dimension must equal the batch"
        nbytes = iq.shape[1]
        n_pcm = self.pcm_count(nbytes)
        if pcm is None:
            pcm = np.zeros((self.batch, n_pcm), np.int16)
        _check(lib().sdr_pipeline_process_host(self._h, iq.ctypes.data, iq.strides[0], nbytes,
                                               pcm.ctypes.data, pcm.strides[0] // 2))
        return pcm

    def process_host_ptr(self, iq_ptr: int, iq_stride: int, nbytes: int, pcm_ptr: int,
                         pcm_stride: int):
        _check(lib().sdr_pipeline_process_host(self._h, iq_ptr, iq_stride, nbytes, pcm_ptr,
                                               pcm_stride))

    def process_device(self, d_iq_ptr: int, iq_stride: int, nbytes: int, d_pcm_ptr: int,
                       pcm_stride: int, stream: int = 0):
        """Raw device pointers (e.g. torch ``tensor.data_ptr()``); only enqueues on ``stream``."""
        _check(lib().sdr_pipeline_process_device(self._h, d_iq_ptr, iq_stride, nbytes, d_pcm_ptr,
                                                 pcm_stride, stream))

    def profile(self, enable=True):
        _check(lib().sdr_pipeline_profile(self._h, 1 if enable else 0))

    def kernel_times(self, reset=False) -> dict:
        """{kernel name: (total_ms, launches)} from CUDA events recorded while profiling."""
        out, i = {}, 0
        while True:
            name = C.create_string_buffer(64)
            ms, cnt = C.c_double(0), C.c_uint64(0)
            rc = lib().sdr_pipeline_kernel_times(self._h, i, name, 64, C.byref(ms), C.byref(cnt),
                                                 1 if reset else 0)
            if rc == 1:
                return out
            _check(rc)
            out[name.value.decode()] = (ms.value, cnt.value)
            i += 1

    def psd(self, name: str):
        """estimatePSD of intermediate `name` of the last call, every capture: (freq[256], psd[batch, 256])."""
        freq = np.zeros(256, np.float32)
        out = np.zeros((self.batch, 256), np.float32)
        _check(lib().sdr_pipeline_psd(self._h, TAP_NAMES.index(name), freq.ctypes.data, out.ctypes.data))
        return freq, out

    def tap(self, name: str, channel: int = 0) -> np.ndarray:
        stage = TAP_NAMES.index(name)
        n = C.c_size_t(0)
        _check(lib().sdr_pipeline_tap(self._h, stage, channel, None, 0, C.byref(n)))
        out = np.zeros(n.value, np.float32)
        _check(lib().sdr_pipeline_tap(self._h, stage, channel, out.ctypes.data, out.size, C.byref(n)))
        return out


# ---- either side of the receiver: PSD diagnostics, de-emphasis, WAV header, channeliser -----------
def estimatePSD(samples, Fs: float, device: int = 0):
    """estimatePSD of the reference (fourier.cpp:44-126) for one row or a [rows, n] array:
    returns (freq[256], psd_dB[rows, 256])."""
    x = np.ascontiguousarray(samples, np.float32)
    one = x.ndim == 1
    if one:
        x = x.reshape(1, -1)
    freq = np.zeros(256, np.float32)
    psd = np.zeros((x.shape[0], 256), np.float32)
    _check(lib().sdr_psd(device, x.ctypes.data, x.shape[0], x.shape[1], x.shape[1], Fs,
                         freq.ctypes.data, psd.ctypes.data))
    return freq, (psd[0] if one else psd)


def wav_header(sample_rate: int, channels: int, n_frames: int) -> bytes:
    buf = (C.c_uint8 * 44)()
    _check(lib().sdr_wav_header(buf, sample_rate, channels, n_frames))
    return bytes(buf)


class Deemphasis:
    """One-pole de-emphasis on the receiver's PCM (int16 [batch, n_frames * channels]), in place."""

    def __init__(self, batch: int, channels: int, Fs: float, tau: float = 75e-6, device: int = 0):
        self.batch, self.channels = batch, channels
        self._h = _vp()
        _check(lib().sdr_deemph_create(device, batch, channels, Fs, tau, C.byref(self._h)))

    def close(self):
        if self._h:
            lib().sdr_deemph_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def reset(self):
        _check(lib().sdr_deemph_reset(self._h))

    def process_host(self, pcm: np.ndarray) -> np.ndarray:
        assert pcm.dtype == np.int16 and pcm.ndim == 2 and pcm.shape[0] == self.batch and pcm.flags.c_contiguous
        _check(lib().sdr_deemph_process_host(self._h, pcm.ctypes.data, pcm.strides[0] // 2,
                                             pcm.shape[1] // self.channels))
        return pcm

    def process_device(self, d_pcm_ptr: int, pcm_stride: int, n_frames: int, stream: int = 0):
        _check(lib().sdr_deemph_process_device(self._h, d_pcm_ptr, pcm_stride, n_frames, stream))


class Channelizer:
    """Polyphase analysis bank in front of the receiver: [n_wide, nbytes] wideband uint8 I/Q at
    n_channels x Fs -> [n_wide * n_channels, nbytes / n_channels] uint8 I/Q at Fs."""

    def __init__(self, n_channels: int, taps_per_branch: int = 16, n_wide: int = 1, device: int = 0, gain: float = 1.0):
        self.M, self.W = n_channels, n_wide
        cfg = ChannelizerConfig(n_channels, taps_per_branch, n_wide, device, gain)
        self._h = _vp()
        _check(lib().sdr_channelizer_create(C.byref(cfg), C.byref(self._h)))

    def close(self):
        if self._h:
            lib().sdr_channelizer_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def reset(self):
        _check(lib().sdr_channelizer_reset(self._h))

    def prototype(self) -> np.ndarray:
        n = C.c_size_t(0)
        _check(lib().sdr_channelizer_prototype(self._h, None, 0, C.byref(n)))
        h = np.zeros(n.value, np.float32)
        _check(lib().sdr_channelizer_prototype(self._h, h.ctypes.data, h.size, C.byref(n)))
        return h

    def process_host(self, wide: np.ndarray) -> np.ndarray:
        wide = np.ascontiguousarray(wide, np.uint8)
        if wide.ndim == 1:
            wide = wide[None, :]
        assert wide.shape[0] == self.W
        out = np.zeros((self.W * self.M, wide.shape[1] // self.M), np.uint8)
        _check(lib().sdr_channelizer_process_host(self._h, wide.ctypes.data, wide.strides[0], wide.shape[1],
                                                  out.ctypes.data, out.strides[0]))
        return out

    def process_device(self, d_wide_ptr: int, wide_stride: int, nbytes_wide: int, d_out_ptr: int, out_stride: int,
                       stream: int = 0):
        _check(lib().sdr_channelizer_process_device(self._h, d_wide_ptr, wide_stride, nbytes_wide, d_out_ptr,
                                                    out_stride, stream))


class MultiPipeline:
    """``batch`` captures spread over several GPUs by one process (sdr_multi_*): contiguous capture
    ranges, one host thread per device, PCM gathered into one host array."""

    def __init__(self, mode=0, channels=1, rf_taps=151, audio_taps=101, stereo_taps=151, batch=1,
                 devices=None, variant=VARIANT_EXACT, max_bytes_per_channel=0):
        cfg = Config(mode, channels, rf_taps, audio_taps, stereo_taps, batch, 0, variant,
                     max_bytes_per_channel)
        self.batch = batch
        if devices is None:
            mc = MultiConfig(cfg, 0, None)
        elif isinstance(devices, int):
            mc = MultiConfig(cfg, devices, None)
        else:
            self._devs = (C.c_int * len(devices))(*devices)
            mc = MultiConfig(cfg, len(devices), self._devs)
        self._h = _vp()
        _check(lib().sdr_multi_create(C.byref(mc), C.byref(self._h)))

    def close(self):
        if self._h:
            lib().sdr_multi_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def layout(self):
        """[(device, first capture), ...]"""
        n = C.c_int(0)
        devs, first = (C.c_int * 64)(), (C.c_int * 64)()
        _check(lib().sdr_multi_layout(self._h, C.byref(n), devs, first, 64))
        return [(devs[i], first[i]) for i in range(n.value)]

    def reset(self):
        _check(lib().sdr_multi_reset(self._h))

    def pcm_count(self, nbytes: int) -> int:
        n = C.c_size_t(0)
        _check(lib().sdr_multi_pcm_count(self._h, nbytes, C.byref(n)))
        return n.value

    def launch_count(self, reset=False) -> int:
        n = C.c_uint64(0)
        _check(lib().sdr_multi_launch_count(self._h, C.byref(n), 1 if reset else 0))
        return n.value

    def process_host(self, iq: np.ndarray, pcm: np.ndarray | None = None) -> np.ndarray:
        iq = np.ascontiguousarray(iq, np.uint8)
        assert iq.ndim == 2 and iq.shape[0] == self.batch, "iq must be [batch, nbytes]"
        nbytes = iq.shape[1]
        if pcm is None:
            pcm = np.zeros((self.batch, self.pcm_count(nbytes)), np.int16)
        _check(lib().sdr_multi_process_host(self._h, iq.ctypes.data, iq.strides[0], nbytes,
                                            pcm.ctypes.data, pcm.strides[0] // 2))
        return pcm

    def process_host_ptr(self, iq_ptr: int, iq_stride: int, nbytes: int, pcm_ptr: int, pcm_stride: int):
        _check(lib().sdr_multi_process_host(self._h, iq_ptr, iq_stride, nbytes, pcm_ptr, pcm_stride))


# ---- RDS chain (model/fmRDS.py:222-276) ----------------------------------------------------------
def rds_design(which: str, mode: int) -> np.ndarray:
    """The model's coefficient sets: 'channel', 'carrier', 'resampler', 'rrc' (float64)."""
    w = RDS_FILTER_NAMES.index(which)
    n = C.c_size_t(0)
    _check(lib().sdr_rds_design(w, mode, None, 0, C.byref(n)))
    h = np.zeros(n.value, np.float64)
    _check(lib().sdr_rds_design(w, mode, h.ctypes.data, h.size, C.byref(n)))
    return h


class Rds:
    """RDS receiver attached to a :class:`Pipeline` (modes 0 and 2): it runs at the end of every
    process call of that pipeline, on the same stream."""

    def __init__(self, pipeline: Pipeline, block_if=0, max_pending_blocks=0, keep_nco=False,
                 cdr_carry=False, pll_form="auto", f32_fir=False):
        self.pipeline = pipeline
        self._h = _vp()
        cfg = RdsConfig(block_if, max_pending_blocks, 1 if keep_nco else 0, 1 if cdr_carry else 0,
                        {"auto": 0, "lane": 1, "warp": 2}[pll_form], 1 if f32_fir else 0)
        _check(lib().sdr_rds_create(pipeline._h, C.byref(cfg), C.byref(self._h)))
        self.info = RdsInfo()
        _check(lib().sdr_rds_info(self._h, C.byref(self.info)))
        if not hasattr(pipeline, "_followers"):
            pipeline._followers = []
        pipeline._followers.append(self)

    def close(self):
        if self._h:
            lib().sdr_rds_destroy(self._h)
            self._h = _vp()
            if self in getattr(self.pipeline, "_followers", []):
                self.pipeline._followers.remove(self)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def pending(self) -> int:
        n = C.c_size_t(0)
        _check(lib().sdr_rds_pending(self._h, C.byref(n)))
        return n.value

    def read(self, channel: int = 0) -> dict:
        """Bit layer of the pending blocks of one capture: cdr_bits, diff_bits (uint8 arrays),
        bit_counts (per block) and offsets (one character per block)."""
        nbits, nblk = C.c_size_t(0), C.c_size_t(0)
        _check(lib().sdr_rds_read(self._h, channel, None, None, 0, C.byref(nbits), None, None, 0,
                                  C.byref(nblk)))
        cdr = np.zeros(max(nbits.value, 1), np.uint8)
        diff = np.zeros(max(nbits.value, 1), np.uint8)
        counts = np.zeros(max(nblk.value, 1), np.int32)
        offs = C.create_string_buffer(max(nblk.value, 1))
        _check(lib().sdr_rds_read(self._h, channel, cdr.ctypes.data, diff.ctypes.data, cdr.size,
                                  C.byref(nbits), counts.ctypes.data, offs, counts.size,
                                  C.byref(nblk)))
        return dict(cdr_bits=cdr[:nbits.value], diff_bits=diff[:nbits.value],
                    bit_counts=counts[:nblk.value], offsets=offs.raw[:nblk.value].decode())

    def discard(self):
        _check(lib().sdr_rds_discard(self._h))

    def tap(self, name: str, channel: int = 0) -> np.ndarray:
        stage = RDS_TAP_NAMES.index(name)
        n = C.c_size_t(0)
        _check(lib().sdr_rds_tap(self._h, stage, channel, None, 0, C.byref(n)))
        out = np.zeros(n.value, np.float64)
        _check(lib().sdr_rds_tap(self._h, stage, channel, out.ctypes.data, out.size, C.byref(n)))
        return out
