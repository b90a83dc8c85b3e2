"""GPU parity tests added in round 2: forked captures (sdr_pipeline_copy_state), the MIXED
variant, the PLL at batches that take its multi-warp launch, the empty-input PLL operator, and
the three-way check reference (oracle/_ref, compiled from the reference's own sources) vs oracle
vs CUDA on the GPU box itself, with the box's libc / CPU recorded next to the result."""
import ctypes
import json
import os
import platform

import numpy as np
import pytest

import orclib
from sdr_b200 import siggen

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def snr_db(got, want):
    got, want = got.astype(np.float64), want.astype(np.float64)
    err = float(np.sum((got - want) ** 2))
    return 400.0 if err == 0 else 10 * np.log10(float(np.sum(want ** 2)) / err)


def test_copy_state_forks_a_capture_in_a_wide_batch(sdr, orc):
    """64 stereo captures; after the first block capture 51 is restarted as a fork of capture 5
    and capture 63 as a fork of capture 0 (rows whose old, wrong copy landed inside OTHER
    captures' history prefixes).  Every capture's second block must equal the oracle fed with the
    same two blocks."""
    B, mode = 64, 0
    bb = siggen.MODES[mode]["block_bytes"]
    iq = siggen.make_batch(B, mode, 2, "stereo", distinct=8)
    for c in range(B):       # make every row distinct
        iq[c] = np.roll(iq[c], 2 * 131 * c)
    first, second = np.ascontiguousarray(iq[:, :bb]), np.ascontiguousarray(iq[:, bb:]).copy()
    forks = {51: 5, 63: 0}
    with sdr.Pipeline(mode=mode, channels=2, batch=B, max_bytes_per_channel=bb) as p:
        p.process_host(first)
        for dst, src in forks.items():
            lib_rc = sdr.lib().sdr_pipeline_copy_state(p._h, dst, src)
            assert lib_rc == 0
        pcm2 = p.process_host(second)
    for c in range(B):
        src = forks.get(c, c)
        want, _ = orc.run_chain(np.concatenate([first[src], second[c]]), mode, 2, keep_taps=False)
        assert np.array_equal(pcm2[c], want[want.size // 2:]), f"capture {c} (history of {src})"


@pytest.mark.parametrize("mode,ch", [(0, 2), (2, 2), (1, 2), (0, 1), (3, 1)])
def test_mixed_variant(sdr, orc, mode, ch):
    """SDR_VARIANT_MIXED: what feeds the PLL is bit-identical to the reference (fm_demod, pilot
    band-pass, NCO); the contracted filters agree to >= 100 dB and the PCM to +-1 LSB."""
    B = 3
    iq = siggen.make_batch(B, mode, 3, "stereo")
    with sdr.Pipeline(mode=mode, channels=ch, batch=B, variant=sdr.VARIANT_MIXED,
                      max_bytes_per_channel=iq.shape[1]) as p:
        p.keep_taps(True)
        pcm = p.process_host(iq)
        exact = ["i_filt", "q_filt", "demod"] + (["carrier_filt", "nco", "allpass"] if ch == 2 else [])
        loose = ["audio_filt"] + (["stereo_filt", "mixer", "stereo_final"] if ch == 2 else [])
        got = {n: [p.tap(n, c) for c in range(B)] for n in exact + loose}
    for c in range(B):
        want_pcm, want = orc.run_chain(iq[c], mode, ch)
        for n in exact:
            assert np.array_equal(bits(got[n][c]), bits(want[n])), f"{n} must stay bit-identical (capture {c})"
        for n in loose:
            assert snr_db(got[n][c], want[n]) >= 100.0, (n, c, snr_db(got[n][c], want[n]))
        d = np.abs(pcm[c].astype(np.int32) - want_pcm.astype(np.int32))
        assert int(d.max()) <= 1, f"PCM differs by {int(d.max())} LSB"


@pytest.mark.parametrize("mode,ch", [(0, 2), (1, 2), (2, 2), (3, 1)])
def test_scalar_form_of_the_exact_fir_kernels(sdr, orc, mode, ch):
    """The exact variant runs its 151-tap FIRs as packed FMUL2 + FFMA2 pairs by default (covered by
    every other exact-variant test); SDR_VARIANT_SCALAR_FIR keeps the FMUL + FADD kernels.  Both
    must give the reference's bits: rf_decim 10 / 5 / 3 front ends and the band-pass pair."""
    B = 3
    iq = siggen.make_batch(B, mode, 2, "stereo")
    names = ["i_filt", "q_filt", "demod", "audio_filt"] + (["stereo_filt", "carrier_filt", "nco"] if ch == 2 else [])
    with sdr.Pipeline(mode=mode, channels=ch, batch=B, variant=sdr.VARIANT_EXACT | sdr.VARIANT_SCALAR_FIR,
                      max_bytes_per_channel=iq.shape[1]) as p:
        p.keep_taps(True)
        pcm = p.process_host(iq)
        got = {n: [p.tap(n, c) for c in range(B)] for n in names}
    for c in range(B):
        want_pcm, want = orc.run_chain(iq[c], mode, ch)
        assert np.array_equal(pcm[c], want_pcm), c
        for n in names:
            assert np.array_equal(bits(got[n][c]), bits(want[n])), (n, c)


def test_pll_multi_warp_launch_at_large_batch(sdr, orc):
    """More than one PLL warp per scheduler (batch > 148*128): k_pll runs four warps per block
    there.  Short captures keep the test small; sampled rows are checked against the oracle."""
    B, nbytes = 19200, 4000
    base = siggen.make_batch(8, 0, 1, "stereo")[:, :nbytes]
    iq = np.ascontiguousarray(np.tile(base, (B // 8, 1)))
    with sdr.Pipeline(mode=0, channels=2, batch=B, max_bytes_per_channel=nbytes) as p:
        pcm = p.process_host(iq)
    for c in (0, 1, 7, 31, 32, 127, 128, 9601, B - 33, B - 1):
        padded = np.full(102400, 128, np.uint8)
        padded[:nbytes] = iq[c]
        want, _ = orc.run_chain(padded, 0, 2, keep_taps=False)
        assert np.array_equal(pcm[c], want[:pcm.shape[1]]), c


def test_pll_operator_with_empty_input(sdr):
    """fmPLL on an empty block: ncoOut = {state[4]}, state untouched (filter.cpp:41-46,73-79)."""
    st = np.array([0.1, 0.2, 0.3, 0.4, 0.5, 6.0], np.float32)
    keep = st.copy()
    out = sdr.fmPLL(np.zeros(0, np.float32), st, 19e3, 240e3, 2.0, 0.0, 0.01)
    assert out.shape == (1,) and out[0] == np.float32(0.5)
    assert np.array_equal(st, keep)


def test_three_way_on_this_host(sdr, orc):
    """Reference (oracle/_ref/libfmref.so: the reference's own filter.cpp + iofunc.cpp compiled
    with its flags) vs the C restatement vs the CUDA path, all on THIS machine.  The PLL parity
    is relative to the host's libm, so libc version and CPU are recorded with the
    result in gpurun_out/three_way.json."""
    ref = orclib.REF()
    if ref is None:
        pytest.skip("oracle/_ref/libfmref.so did not travel to this host")
    libc = ctypes.CDLL("libc.so.6")
    libc.gnu_get_libc_version.restype = ctypes.c_char_p
    flags = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                flags = " ".join(f for f in line.split(":", 1)[1].split() if f in ("fma", "avx2", "avx512f", "sse4_2"))
                break
    except OSError:
        pass
    record = {"glibc": libc.gnu_get_libc_version().decode(), "machine": platform.machine(),
              "cpu_flags": flags, "cases": []}
    for mode, ch, kind in [(0, 2, "stereo"), (2, 2, "stereo"), (1, 2, "rds"), (3, 2, "stereo"), (0, 1, "stereo"), (2, 1, "mono")]:
        iq = siggen.make_batch(2, mode, 3, kind)
        with sdr.Pipeline(mode=mode, channels=ch, batch=2, max_bytes_per_channel=iq.shape[1]) as p:
            p.keep_taps(True)
            pcm = p.process_host(iq)
            nco = [p.tap("nco", c) for c in range(2)] if ch == 2 else None
        for c in range(2):
            r_pcm, r_taps = ref.run_chain(iq[c], mode, ch)
            o_pcm, _ = orc.run_chain(iq[c], mode, ch, keep_taps=False)
            assert np.array_equal(r_pcm, o_pcm), "oracle restatement differs from the compiled reference on this host"
            assert np.array_equal(pcm[c], r_pcm), f"CUDA PCM differs from the compiled reference (mode {mode}, ch {ch})"
            if nco is not None:
                assert np.array_equal(bits(nco[c]), bits(r_taps["nco"])), "NCO differs from the reference's fmPLL"
        record["cases"].append({"mode": mode, "channels": ch, "captures": 2, "pcm_values": int(pcm.size),
                                "bit_identical": True})
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "three_way.json"), "w") as f:
        json.dump(record, f, indent=1)
    print("three-way parity on this host:", json.dumps(record))


# ---- one process, several devices (sdr_multi_*) ------------------------------------------------
@pytest.mark.parametrize("mode,ch,variant", [(0, 2, "exact"), (2, 1, "fast")])
def test_multi_device_equals_single_pipeline(sdr, orc, mode, ch, variant):
    """The whole batch through sdr_multi_process_host (every device of the box, contiguous capture
    ranges, PCM gathered into one host array) == one pipeline on device 0 == the oracle; the same
    through two consecutive calls (carried state per device)."""
    v = sdr.VARIANT_FAST if variant == "fast" else sdr.VARIANT_EXACT
    B = 13   # not a multiple of any device count
    iq = siggen.make_batch(B, mode, 2, "stereo", distinct=5)
    half = iq.shape[1] // 2
    with sdr.Pipeline(mode=mode, channels=ch, batch=B, variant=v, max_bytes_per_channel=iq.shape[1]) as p:
        one = p.process_host(iq)
    n_dev = sdr.device_count()
    for devices in sorted({1, min(2, n_dev), n_dev}):
        with sdr.MultiPipeline(mode=mode, channels=ch, batch=B, devices=devices, variant=v,
                               max_bytes_per_channel=iq.shape[1]) as m:
            lay = m.layout()
            assert len(lay) == devices and lay[0] == (0, 0)
            whole = m.process_host(iq)
            m.reset()
            parts = [m.process_host(np.ascontiguousarray(iq[:, :half])),
                     m.process_host(np.ascontiguousarray(iq[:, half:]))]
            assert m.launch_count() > 0
        assert np.array_equal(whole, one), f"{devices} device(s): differs from the single pipeline"
        assert np.array_equal(np.concatenate(parts, axis=1), one), f"{devices} device(s): two calls differ"
    for c in (0, 6, B - 1):
        want, _ = orc.run_chain(iq[c], mode, ch, keep_taps=False)
        d = np.abs(one[c].astype(np.int32) - want.astype(np.int32)).max()
        assert d <= (1 if variant == "fast" else 0)


def test_multi_device_needs_at_least_two_gpus_for_the_split(sdr):
    """On a multi-GPU box the ranges really land on different devices."""
    if sdr.device_count() < 2:
        pytest.skip("one GPU on this box")
    with sdr.MultiPipeline(mode=0, channels=1, batch=10, devices=2, max_bytes_per_channel=102400) as m:
        assert m.layout() == [(0, 0), (1, 5)]


def test_project_cli_batch(orc):
    """`sdr_project 0 2 --batch 3 --blocks 2`: stdin carries per call 3 chunks of 2 blocks (one per
    capture), stdout the 3 PCM chunks in the same order; the trailing incomplete round is dropped."""
    import subprocess
    pkg = os.path.join(ROOT, "software-defined-radio_b200")
    B, blocks, calls, mode = 3, 2, 2, 0
    bb = siggen.MODES[mode]["block_bytes"]
    iq = siggen.make_batch(B, mode, blocks * calls, "stereo")
    stream = b"".join(iq[c, k * blocks * bb:(k + 1) * blocks * bb].tobytes() for k in range(calls) for c in range(B))
    r = subprocess.run([os.path.join(pkg, "sdr_project"), "0", "2", "--batch", str(B), "--blocks", str(blocks)],
                       input=stream + b"\x80" * 5000, capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    got = np.frombuffer(r.stdout, dtype=np.int16).reshape(calls, B, -1)
    for c in range(B):
        want, _ = orc.run_chain(iq[c], mode, 2, keep_taps=False)
        assert np.array_equal(np.concatenate([got[k, c] for k in range(calls)]), want), c


# ---- INTEGRATION.md level 1: the reference's own project.cpp on top of the drop-in filter.h --------
def test_level1_reference_project_cpp_runs_on_the_dropin():
    """oracle/_ref/project_level1 is the reference's UNMODIFIED src/project.cpp (+ iofunc, fourier,
    genfunc, logfunc; not its filter.cpp) compiled with include/dropin/ in front of its own
    include/ and linked -lsdr_filter -lsdr_b200 (oracle/Makefile, i.e. src/Makefile:3-10 with one
    -I and two -l added); oracle/_ref/project_ref is the reference's own build.  Both get the
    same two stereo blocks: same exit status (the reference's exit(1) at end of input,
    project.cpp:83-93) and the same report on stdout once the timing figures are masked."""
    import re
    import subprocess
    l1 = os.path.join(ROOT, "oracle", "_ref", "project_level1")
    rf = os.path.join(ROOT, "oracle", "_ref", "project_ref")
    if not (os.path.exists(l1) and os.path.exists(rf)):
        pytest.skip("oracle/_ref/project_level1 was not built (needs /root/reference at build time)")
    iq = siggen.make_capture(3, 0, 2, "stereo").tobytes()

    def run(exe):
        r = subprocess.run([exe, "0", "2"], input=iq, capture_output=True, timeout=300)
        text = re.sub(r"[-+]?\d+(\.\d+)?(e[-+]?\d+)?", "#", r.stdout.decode())
        # the producer's exit(1) races the consumer: keep what does not depend on it
        head = text.split("___________________Read block")[0]
        # (two threads write to stdout unsynchronised: compare the producer's closing phrases in order,
        # not whole lines, which the consumer's per-block report may cut into)
        tail = re.findall(r"End of input stream reached|I Filter FINAL|Q Filter FINAL|fmDemod FINAL|Program ran for", text)
        return r.returncode, head, tail, r.stderr.decode()

    rc1, head1, tail1, err1 = run(l1)
    rc0, head0, tail0, _ = run(rf)
    assert rc0 == 1 and rc1 == 1, (rc0, rc1, err1[-1500:])
    assert head1 == head0 and tail1 == tail0
    assert "End of input stream reached" in "".join(tail1)


# ---- tensor-core resampler (csrc/resample_tc.cuh) -------------------------------------------------
@pytest.mark.parametrize("mode", [2, 3])
def test_tc_resampler_wide_batch_and_streaming(sdr, orc, mode):
    """FAST mono, modes 2/3: 130 captures (a second, almost empty 128-capture tile), the capture cut
    into three uneven calls (the fp16 planes' history is carried) must give the same bits as one
    call; audio_filt >= 100 dB and PCM within +-1 LSB of the reference on sampled captures."""
    B = 130
    iq = siggen.make_batch(B, mode, 3, "stereo", distinct=6)
    gran = sdr.mode_info(mode, 1).granule_bytes
    n = iq.shape[1]
    cuts = [0, gran * 2, gran * 9, n]
    with sdr.Pipeline(mode=mode, channels=1, batch=B, variant=sdr.VARIANT_FAST, max_bytes_per_channel=n) as p:
        p.keep_taps(True)
        p.profile(True)
        whole = p.process_host(iq)
        audio = {c: p.tap("audio_filt", c) for c in (0, 5, 127, 128, 129)}
        assert "k_audio_resample" in p.kernel_times()
        # both custom-rate modes run the tensor-core resampler (the front end then writes the fp16 planes)
        p.keep_taps(False)
        p.reset()
        parts = [p.process_host(np.ascontiguousarray(iq[:, a:b])) for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.array_equal(np.concatenate(parts, axis=1), whole), "result depends on how the capture is cut into calls"
    for c, got in audio.items():
        want_pcm, want = orc.run_chain(iq[c], mode, 1)
        assert snr_db(got, want["audio_filt"]) >= 100.0, (c, snr_db(got, want["audio_filt"]))
        d = np.abs(whole[c].astype(np.int32) - want_pcm.astype(np.int32))
        assert int(d.max()) <= 1, f"capture {c}: PCM differs by {int(d.max())} LSB"


@pytest.mark.parametrize("mode", [2, 0])
def test_many_short_device_calls_back_to_back(sdr, orc, mode):
    """The fast mono path's kernels are launched with programmatic dependence (each may be scheduled
    before its predecessor finishes and waits for it itself).  Dozens of granule-sized calls enqueued back
    to back on one stream, nothing but the launches in between -- every kernel is a few microseconds, so
    the overlap windows are as large as they get relative to the work -- must give the bits of one call."""
    import torch
    B = 257
    gran = sdr.mode_info(mode, 1).granule_bytes
    per = gran * (1 if mode == 2 else 16)
    bb = siggen.MODES[mode]["block_bytes"]
    calls = (6 if mode == 2 else 1) * bb // per          # whole reference blocks (the oracle's unit): 42 / 64 calls
    n = per * calls
    src = siggen.make_batch(6, mode, n // bb, "stereo")[:, :n]
    iq = np.ascontiguousarray(np.tile(src, (B // 6 + 1, 1))[:B])
    for c in range(B):
        iq[c] = np.roll(iq[c], 2 * 17 * c)
    d_iq = torch.from_numpy(iq).cuda()
    s = torch.cuda.current_stream().cuda_stream
    with sdr.Pipeline(mode=mode, channels=1, batch=B, variant=sdr.VARIANT_FAST, max_bytes_per_channel=n) as p:
        n_pcm = p.pcm_count(n)
        one = torch.zeros((B, n_pcm), dtype=torch.int16, device="cuda")
        p.process_device(d_iq.data_ptr(), d_iq.stride(0), n, one.data_ptr(), one.stride(0), s)
        torch.cuda.synchronize()
        p.reset()
        many = torch.zeros_like(one)
        k = p.pcm_count(per)
        for i in range(calls):
            p.process_device(d_iq.data_ptr() + i * per, d_iq.stride(0), per, many.data_ptr() + 2 * i * k, many.stride(0), s)
        torch.cuda.synchronize()
    assert torch.equal(one, many)
    want, _ = orc.run_chain(iq[B - 1], mode, 1, keep_taps=False)
    d = np.abs(one[B - 1].cpu().numpy().astype(np.int32) - want[:n_pcm].astype(np.int32))
    assert int(d.max()) <= 1
