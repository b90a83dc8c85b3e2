"""GPU tests of the stages either side of the hot path (csrc/aux.cu): the PSD diagnostics op against
the oracle's restatement of estimatePSD and the golden vectors made from the compiled reference,
de-emphasis bit for bit against its oracle, the channeliser against its direct-form oracle and end
to end (one wideband capture -> channels in device memory -> the batched receiver -> PCM)."""
import os
import subprocess

import numpy as np
import pytest

import auxlib
from sdr_b200 import siggen

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def close_db(got, want):
    """The O(N^2) float DFT of the reference leaves bins far below the strongest one to rounding
    (a last-bit difference in one twiddle moves a -150 dB bin by 0.01 dB): 1e-3 dB within 80 dB of
    the peak, 0.1 dB below that."""
    d = np.abs(got - want)
    strong = want >= want.max() - 80.0
    return float(d[strong].max()) <= 1e-3 and float(d.max()) <= 0.1


def test_psd_op_matches_reference_golden_vectors(sdr):
    g = np.load(os.path.join(GOLD, "psd.npz"))
    for name in ("audio48k", "if240k", "one_segment"):
        freq, psd = sdr.estimatePSD(g[f"{name}_x"], float(g[f"{name}_fs"]))
        assert np.array_equal(freq, g[f"{name}_freq"]), name
        assert close_db(psd, g[f"{name}_psd"]), name


def test_psd_op_batched_rows_match_oracle(sdr):
    rng = np.random.default_rng(11)
    x = (0.25 * rng.standard_normal((70, 512 * 5 + 99))).astype(np.float32)   # 70 rows: more than one CTA of segments
    x[3] = 0.5 * np.sin(2 * np.pi * 0.1 * np.arange(x.shape[1])).astype(np.float32)
    freq, psd = sdr.estimatePSD(x, 240000.0)
    for r in (0, 3, 33, 69):
        fo, po = auxlib.psd(x[r], 240000.0)
        assert np.array_equal(freq, fo)
        assert close_db(psd[r], po), r
    with pytest.raises(sdr.SdrError):
        sdr.estimatePSD(np.zeros(100, np.float32), 48000.0)   # shorter than one segment


def test_pipeline_psd_of_intermediates(sdr, orc):
    iq = siggen.make_batch(3, 0, 2, "stereo")
    with sdr.Pipeline(mode=0, channels=2, batch=3, max_bytes_per_channel=iq.shape[1]) as p:
        p.keep_taps(True)
        p.process_host(iq)
        got = {name: p.psd(name) for name in ("demod", "audio_filt", "carrier_filt", "stereo_final")}
    for c in range(3):
        _, taps = orc.run_chain(iq[c], 0, 2)
        for name, fs in (("demod", 240000.0), ("audio_filt", 48000.0), ("carrier_filt", 240000.0), ("stereo_final", 48000.0)):
            fo, po = auxlib.psd(taps[name], fs)
            assert np.array_equal(got[name][0], fo)
            assert close_db(got[name][1][c], po), (name, c)
    # the pilot shows up where it should: the strongest bin of the pilot band-pass output is 19 kHz
    f, ps = got["carrier_filt"]
    assert abs(f[int(np.argmax(ps[0]))] - 19000.0) <= 240000.0 / 512


@pytest.mark.parametrize("channels,fs", [(1, 48000.0), (2, 44100.0)])
def test_deemphasis_bit_identical_to_its_oracle(sdr, channels, fs):
    rng = np.random.default_rng(5)
    B, frames = 37, 3000
    pcm = rng.integers(-20000, 20000, size=(B, frames * channels)).astype(np.int16)
    want = pcm.copy()
    for b in range(B):
        st = np.zeros(channels, np.float32)
        auxlib.deemphasis(want[b], channels, fs, 75e-6, st)
    with sdr.Deemphasis(B, channels, fs, 75e-6) as d:
        one = d.process_host(pcm.copy())
        d.reset()
        cut = 1001 * channels
        two = np.concatenate([d.process_host(np.ascontiguousarray(pcm[:, :cut])),
                              d.process_host(np.ascontiguousarray(pcm[:, cut:]))], axis=1)
    assert np.array_equal(one, want)
    assert np.array_equal(two, want), "carried state"


@pytest.mark.parametrize("M,T", [(2, 8), (4, 16), (8, 12), (16, 6)])
def test_channelizer_matches_direct_form_oracle(sdr, M, T):
    rng = np.random.default_rng(M)
    n_pairs = 2500 * M
    wide = np.stack([siggen.make_wideband(M, 2500)[0],
                     rng.integers(0, 256, size=2 * n_pairs).astype(np.uint8)])   # stations / full-scale noise
    with sdr.Channelizer(M, T, n_wide=2, gain=1.5) as ch:
        h = ch.prototype()
        assert h.size == M * T and np.array_equal(h, sdr.impulseResponseLPF(float(M), 0.4, M * T))
        got = ch.process_host(wide)
        ch.reset()
        cut = 2 * M * 700
        parts = np.concatenate([ch.process_host(np.ascontiguousarray(wide[:, :cut])),
                                ch.process_host(np.ascontiguousarray(wide[:, cut:]))], axis=1)
    assert np.array_equal(parts, got), "filter history carried across calls"
    for w in range(2):
        want = auxlib.channelize(wide[w], M, h, 1.5)
        d = np.abs(got[w * M:(w + 1) * M].astype(np.int32) - want.astype(np.int32))
        assert int(d.max()) <= 1, (w, int(d.max()))                      # float vs double at rounding boundaries
        assert float(np.mean(d != 0)) < 0.01, float(np.mean(d != 0))


@pytest.mark.parametrize("M,T,n_inst", [(16, 5, 300), (16, 8, 3), (8, 7, 513), (4, 3, 2049), (2, 64, 40)])
def test_channelizer_ragged_shapes(sdr, M, T, n_inst):
    """Edges of the throughput kernel's tiling: tap counts that are not a multiple of its 4-tap chunk,
    fewer output instants than taps, one instant past a tile (256 / 512 / 1024 / 2048 instants per tile
    for M = 16 / 8 / 4 / 2), the longest branch filter, and a capture cut into three uneven calls."""
    rng = np.random.default_rng(100 * M + T)
    wide = rng.integers(0, 256, size=(3, 2 * M * n_inst)).astype(np.uint8)
    wide[1] = 128                                             # silence
    wide[2, ::2] = 255                                        # clipped I
    with sdr.Channelizer(M, T, n_wide=3, gain=0.9) as ch:
        h = ch.prototype()
        got = ch.process_host(wide)
        ch.reset()
        a, b = 2 * M * (n_inst // 3), 2 * M * (n_inst // 3 + max(1, n_inst // 2))
        cuts = [c for c in (0, a, min(b, wide.shape[1]), wide.shape[1])]
        parts = [ch.process_host(np.ascontiguousarray(wide[:, cuts[k]:cuts[k + 1]]))
                 for k in range(3) if cuts[k + 1] > cuts[k]]
    assert np.array_equal(np.concatenate(parts, axis=1), got), "history carried across uneven calls"
    for w in range(3):
        want = auxlib.channelize(wide[w], M, h, 0.9)
        d = np.abs(got[w * M:(w + 1) * M].astype(np.int32) - want.astype(np.int32))
        assert int(d.max()) <= 1, (w, int(d.max()))
    assert np.all(got[M:2 * M] == 128), "silence in, silence out"


def test_wideband_capture_to_pcm_on_the_device(sdr):
    """eight stations in one 19.2 MS/s capture, channelised in device memory and handed
    to the batched receiver without crossing PCIe again; every channel's PCM carries its own tone."""
    import torch
    M, blocks = 8, 4
    n_ch_pairs = blocks * 51200
    wide, amp = siggen.make_wideband(M, n_ch_pairs)
    d_wide = torch.from_numpy(wide).cuda()
    d_chan = torch.empty((M, 2 * n_ch_pairs), dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    with sdr.Channelizer(M, 16, n_wide=1, gain=0.6 / amp) as ch, \
            sdr.Pipeline(mode=0, channels=1, batch=M, variant=sdr.VARIANT_FAST, max_bytes_per_channel=2 * n_ch_pairs) as p:
        d_pcm = torch.zeros((M, p.pcm_count(2 * n_ch_pairs)), dtype=torch.int16, device="cuda")
        ch.process_device(d_wide.data_ptr(), wide.size, wide.size, d_chan.data_ptr(), d_chan.stride(0), s)
        p.process_device(d_chan.data_ptr(), d_chan.stride(0), 2 * n_ch_pairs, d_pcm.data_ptr(), d_pcm.stride(0), s)
        torch.cuda.synchronize()
        pcm = d_pcm.cpu().numpy().astype(np.float64)
    for c in range(M):
        x = pcm[c, 512:]
        spec = np.abs(np.fft.rfft(x * np.hanning(x.size)))
        peak = np.argmax(spec[5:]) + 5
        assert abs(peak * 48000.0 / x.size - (1000.0 + 400.0 * c)) < 25.0, c
        assert x.std() > 500


def test_project_cli_wav_and_deemphasis(tmp_path, orc):
    """`sdr_project 0 2 --wav FILE --deemphasis 75`: same PCM as stdout carries, de-emphasised, in a
    RIFF/WAVE file whose header matches what was written."""
    import struct
    iq = siggen.make_capture(9, 0, 3, "stereo")
    wav = tmp_path / "out.wav"
    r = subprocess.run([os.path.join(ROOT, "software-defined-radio_b200", "sdr_project"), "0", "2", "--wav", str(wav),
                        "--deemphasis", "75"], input=iq.tobytes(), capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-1500:]
    want, _ = orc.run_chain(iq, 0, 2, keep_taps=False)
    st = np.zeros(2, np.float32)
    auxlib.deemphasis(want, 2, 48000.0, 75e-6, st)
    assert np.array_equal(np.frombuffer(r.stdout, dtype=np.int16), want)
    raw = wav.read_bytes()
    data, = struct.unpack("<I", raw[40:44])
    assert raw[:4] == b"RIFF" and data == want.size * 2 and len(raw) == 44 + data
    assert np.array_equal(np.frombuffer(raw[44:], dtype=np.int16), want)
