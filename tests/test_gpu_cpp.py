"""GPU tests of the C++ host side: the drop-in filter.h shim (C++ caller -> C ABI -> CUDA) and
the `sdr_project` executable (stdin uint8 I/Q -> stdout int16 PCM), against the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

from sdr_b200 import siggen

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "software-defined-radio_b200")


def test_dropin_filter_h_from_cpp(orc, tmp_path):
    exe = tmp_path / "shim_parity"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I",
                    os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "cpp", "shim_parity.cpp"),
                    "-o", str(exe), "-L", PKG, "-lsdr_filter", "-lsdr_b200", "-L",
                    os.path.join(ROOT, "oracle"), "-lfm_oracle",
                    f"-Wl,-rpath,{PKG}", f"-Wl,-rpath,{os.path.join(ROOT, 'oracle')}"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("argv,mode,ch", [([], 0, 1), (["1"], 1, 1), (["0", "2"], 0, 2), (["2", "2"], 2, 2),
                                          (["3", "1", "--blocks", "3"], 3, 1)])
def test_project_cli_stdin_to_stdout(orc, argv, mode, ch):
    """The reference's process contract (project.cpp:385-500): raw I/Q in, PCM out, trailing
    partial block dropped, everything drained at EOF, exit status 0."""
    iq = siggen.make_capture(40 + mode, mode, 5, "stereo")
    data = iq.tobytes() + b"\x80" * 1001  # partial trailing block must be ignored
    r = subprocess.run([os.path.join(PKG, "sdr_project"), *argv], input=data, capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    got = np.frombuffer(r.stdout, dtype=np.int16)
    want, _ = orc.run_chain(iq, mode, ch, keep_taps=False)
    assert np.array_equal(got, want)


def test_project_cli_rejects_bad_mode():
    r = subprocess.run([os.path.join(PKG, "sdr_project"), "7"], input=b"", capture_output=True, timeout=60)
    assert r.returncode == 1 and b"Wrong mode" in r.stderr


@pytest.mark.parametrize("mode,n_ref", [(0, 30), (2, 24)])
def test_project_cli_rds(orc, tmp_path, mode, n_ref):
    """`sdr_project <mode> 1 --rds FILE`: PCM on stdout as before, one line per RDS block in FILE
    with the frame synchroniser's result and the differentially decoded bits -- identical to
    the oracle's restatement of model/fmRDS.py:222-276 on the same fm_demod."""
    import orclib
    iq = siggen.make_capture(5, mode, n_ref, "rds_groups")
    out = tmp_path / "rds.txt"
    r = subprocess.run([os.path.join(PKG, "sdr_project"), str(mode), "1", "--rds", str(out)],
                       input=iq.tobytes(), capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    want_pcm, taps = orc.run_chain(iq, mode, 1)
    assert np.array_equal(np.frombuffer(r.stdout, dtype=np.int16), want_pcm)
    fm = taps["demod"].astype(np.float64)
    want = orclib.RDS().run_chain(fm[:fm.size // 9600 * 9600], mode, 9600, keep=())
    lines = out.read_text().splitlines()
    assert len(lines) == len(want["diff_bits"])
    for i, line in enumerate(lines):
        idx, off, *bits = line.split(" ")
        assert int(idx) == i
        assert off == (want["offsets"][i] if want["offsets"][i] != " " else "-")
        assert (bits[0] if bits else "") == "".join(str(int(b)) for b in want["diff_bits"][i])
