"""GPU tests of SDR_VARIANT_FAST (tensor-core RF front end, mono, all four modes).

Two bars: (1) against an exact integer model of what the kernel is specified to compute (fixed-
point taps, int64 sums, one rounding to float) the I/Q outputs must be within ONE ULP, and exactly
zero where the model is zero (the sums are exact; only the final int -> float recombination
rounds, three times instead of once); (2) against
the reference oracle the task's tolerance applies: float intermediates >= 100 dB SNR (1e-5), PCM
within +-1 LSB."""
import numpy as np
import pytest

from sdr_b200 import siggen

pytestmark = pytest.mark.gpu


def snr_db(ref, got):
    ref = ref.astype(np.float64); got = got.astype(np.float64)
    err = np.sum((ref - got) ** 2)
    return np.inf if err == 0 else 10 * np.log10(np.sum(ref ** 2) / err)


def fixed_point_model(iq, h, decim):
    """I/Q of the tensor-core front end, evaluated exactly on the host."""
    hmax = float(np.max(np.abs(h)))
    S = 0
    while S < 60 and np.ldexp(hmax, S + 1) < 1073741823.0:   # 31-bit fixed point: four base-256 digits
        S += 1
    hq = np.rint(np.ldexp(h.astype(np.float64), S)).astype(np.int64)
    out = []
    for comp in (0, 1):
        x = iq[comp::2].astype(np.int64) - 128
        full = np.convolve(x, hq)[: x.size]          # y[m] = sum_t hq[t] x[m-t], zero history
        v = full[::decim]
        out.append((v.astype(np.float32) * np.float32(np.ldexp(1.0, -(S + 7)))).astype(np.float32))
    return out


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("taps", [(151, 101), (13, 13), (64, 101)])
def test_fast_front_end(sdr, orc, mode, taps):
    B = 3
    iq = siggen.make_batch(B, mode, 3, "stereo")
    nbytes = iq.shape[1]
    with sdr.Pipeline(mode=mode, channels=1, rf_taps=taps[0], audio_taps=taps[1], batch=B,
                      variant=sdr.VARIANT_FAST, max_bytes_per_channel=nbytes) as p:
        p.keep_taps(True)
        p.profile(True)
        pcm = p.process_host(iq)
        got = {n: [p.tap(n, c) for c in range(B)] for n in ("i_filt", "q_filt", "demod", "audio_filt")}
        launched = p.kernel_times()
        assert "k_rf_demod_tc" in launched and "k_rf_demod" not in launched, \
            f"the fast variant must run the tensor-core front end, ran {sorted(launched)}"
    rf_Fs, decim = sdr.mode_info(mode).rf_Fs, sdr.mode_info(mode).rf_decim
    h = sdr.impulseResponseLPF(rf_Fs, 100000, taps[0])
    for c in range(B):
        mi, mq = fixed_point_model(iq[c], h, decim)
        for name, model in (("i_filt", mi), ("q_filt", mq)):
            err = np.abs(got[name][c].astype(np.float64) - model.astype(np.float64))
            assert np.all(err <= np.spacing(np.abs(model))), f"{name} is more than one ulp from the integer model"
            assert not got[name][c][model == 0].any(), f"{name}: exact zeros of the model are not zero"
        want_pcm, want = orc.run_chain(iq[c], mode, 1, taps[0], taps[1], 151)
        for name in ("i_filt", "q_filt", "demod", "audio_filt"):
            s = snr_db(want[name], got[name][c])
            assert s >= 100.0, f"{name}: {s:.1f} dB"
        d = np.abs(pcm[c].astype(np.int32) - want_pcm.astype(np.int32))
        assert d.max() <= 1, f"PCM differs by {d.max()} LSB"


@pytest.mark.parametrize("mode,cuts", [(0, [0, 70000, 70100, 150000]), (1, [0, 600, 61440, 100020]),
                                       (3, [0, 19200, 134400, 172800])])
def test_fast_streaming_and_wide_batch(sdr, orc, mode, cuts):
    """Carried raw history / predecessor output across calls and segments; 200 captures."""
    B = 200
    iq = siggen.make_batch(B, mode, 2, "stereo", distinct=5)
    nbytes = iq.shape[1]
    with sdr.Pipeline(mode=mode, channels=1, batch=B, variant=sdr.VARIANT_FAST, max_bytes_per_channel=nbytes) as p:
        one = p.process_host(iq)
        p.reset()
        cuts = cuts + [nbytes]
        parts = [p.process_host(np.ascontiguousarray(iq[:, a:b])) for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.array_equal(one, np.concatenate(parts, axis=1)), "chunked calls differ from one call"
    for c in (0, 4, 5, 199):
        want, _ = orc.run_chain(iq[c], mode, 1, keep_taps=False)
        assert np.abs(one[c].astype(np.int32) - want.astype(np.int32)).max() <= 1


def test_fast_many_work_items_are_consistent(sdr):
    """More (capture, segment) work items than resident CTAs: every CTA pipelines several items
    back to back.  The result must not depend on how the work was cut: rows holding the same
    capture are identical, and equal to a 3-capture pipeline (fewer items than CTAs)."""
    B, mode = 600, 2
    base = siggen.make_batch(3, mode, 4, "mono")
    iq = np.ascontiguousarray(base[np.arange(B) % 3])
    nbytes = iq.shape[1]
    with sdr.Pipeline(mode=mode, channels=1, batch=B, variant=sdr.VARIANT_FAST, max_bytes_per_channel=nbytes) as p:
        wide = p.process_host(iq)
    with sdr.Pipeline(mode=mode, channels=1, batch=3, variant=sdr.VARIANT_FAST, max_bytes_per_channel=nbytes) as p:
        small = p.process_host(base)
    assert np.array_equal(wide, small[np.arange(B) % 3])


def test_fast_silence_is_exactly_zero(sdr):
    iq = np.full((2, 102400), 128, np.uint8)
    with sdr.Pipeline(mode=0, channels=1, batch=2, variant=sdr.VARIANT_FAST, max_bytes_per_channel=102400) as p:
        assert not p.process_host(iq).any()


def test_fast_refuses_stereo(sdr):
    for kw in (dict(mode=0, channels=2), dict(mode=3, channels=2), dict(mode=0, channels=1, rf_taps=201)):
        with pytest.raises(sdr.SdrError) as e:
            sdr.Pipeline(variant=sdr.VARIANT_FAST, **kw)
        assert e.value.code == -1
