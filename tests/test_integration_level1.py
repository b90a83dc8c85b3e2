"""INTEGRATION.md level 1, CPU side: the reference's unmodified src/project.cpp compiles against
include/dropin/filter.h and links with libsdr_filter.so + libsdr_b200.so (no GPU needed for that).
Without a GPU the resulting binary must fail loudly -- there is no CPU fallback -- instead of
producing output.  The GPU side of the same recipe is tests/test_gpu_round2.py."""
import os
import subprocess

import numpy as np
import pytest

import sdr_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference/src"


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="/root/reference is not on this host")
def test_reference_project_cpp_builds_against_the_dropin(tmp_path):
    pkg = os.path.join(ROOT, "software-defined-radio_b200")
    exe = tmp_path / "project_level1"
    srcs = [os.path.join(REF_SRC, f) for f in ("project.cpp", "iofunc.cpp", "fourier.cpp", "genfunc.cpp", "logfunc.cpp")]
    r = subprocess.run(["g++", "-O3", "-pthread", "-I", os.path.join(ROOT, "include", "dropin"),
                        "-I", "/root/reference/include", *srcs, "-o", str(exe), "-L", pkg, "-lsdr_filter",
                        "-lsdr_b200", f"-Wl,-rpath,{pkg}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    # every filter.h symbol project.cpp uses resolves to the shim, none to a reference object
    nm = subprocess.run(["nm", "-C", "--undefined-only", str(exe)], capture_output=True, text=True).stdout
    for fn in ("impulseResponseLPF", "convolveBlockFastFIR", "fmDemod", "fmPLL", "bandPass", "allPass"):
        assert fn in nm, f"{fn} is not an undefined (shim-provided) symbol of the level-1 binary"
    if sdr_b200.device_count() == 0:
        data = np.full(2 * 102400, 128, np.uint8).tobytes()
        r = subprocess.run([str(exe), "0", "2"], input=data, capture_output=True, timeout=120)
        # the shim throws std::runtime_error (no sm_100 device): the process aborts, never "succeeds"
        assert r.returncode not in (0, 1), "without a GPU the level-1 binary must fail, not fall back"
        assert b"no CUDA device" in r.stderr or b"sm_100" in r.stderr or b"terminate" in r.stderr
