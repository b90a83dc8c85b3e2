"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/sdr_b200.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import sdr_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sdr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sdr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(sdr_b200.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sdr_b200.h but not exported"
    # and the Python binding covers the same set
    assert set(names) == set(sdr_b200.ABI), set(names) ^ set(sdr_b200.ABI)


def test_mode_table_and_granules():
    # project.cpp:424-427 / :55-57
    want = {0: (2400000, 240000, 48000, 10, 5, 1, 102400, 100, 1),
            1: (1440000, 288000, 48000, 5, 6, 1, 61440, 60, 1),
            2: (2400000, 240000, 44100, 10, 800, 147, 112000, 16000, 147),
            3: (960000, 320000, 44100, 3, 3200, 441, 134400, 19200, 441)}
    for mode, w in want.items():
        mi = sdr_b200.mode_info(mode, 1)
        got = (mi.rf_Fs, mi.if_Fs, mi.audio_Fs, mi.rf_decim, mi.audio_decim, mi.audio_upsamp,
               mi.block_bytes, mi.granule_bytes, mi.pcm_per_granule)
        assert got == w
        assert mi.block_bytes % mi.granule_bytes == 0
        assert sdr_b200.mode_info(mode, 2).pcm_per_granule == 2 * w[-1]
    with pytest.raises(sdr_b200.SdrError):
        sdr_b200.mode_info(4, 1)
    with pytest.raises(sdr_b200.SdrError):
        sdr_b200.mode_info(0, 3)


def test_bad_config_is_rejected_before_touching_the_gpu():
    for kw in (dict(mode=7), dict(channels=3), dict(batch=0), dict(rf_taps=1),
               dict(mode=3, audio_taps=151)):  # 151*441 > 65535 (filter.h:24 unsigned short)
        with pytest.raises(sdr_b200.SdrError) as e:
            sdr_b200.Pipeline(**kw)
        assert e.value.code == -1, kw


@pytest.mark.skipif(sdr_b200.device_count() > 0, reason="host has a GPU")
def test_no_cpu_fallback():
    with pytest.raises(sdr_b200.SdrError) as e:
        sdr_b200.Pipeline()
    assert e.value.code == -2
    with pytest.raises(sdr_b200.SdrError) as e:
        sdr_b200.convolveBlockFIR(np.zeros(8, np.float32), np.ones(3, np.float32), np.zeros(2, np.float32))
    assert e.value.code == -2
