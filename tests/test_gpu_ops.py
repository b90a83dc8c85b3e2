"""GPU parity tests, operator by operator, through the C ABI (host-pointer entry points) against
the CPU oracle.  Everything is compared bit for bit: the CUDA path performs the reference's float
operations in the reference's order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same(a, b):
    return np.array_equal(bits(a), bits(b))


def test_device_is_present(sdr):
    assert sdr.device_count() >= 1, "GPU tests need an sm_100 device; there is no CPU fallback"


@pytest.mark.parametrize("nh,decim,nx", [(13, 1, 1000), (151, 1, 5120), (151, 10, 51200),
                                         (101, 5, 5120), (37, 3, 3000), (2, 1, 17), (151, 5, 30720)])
def test_fir_block_and_decim(sdr, orc, nh, decim, nx):
    rng = np.random.default_rng(nh * 1000 + decim)
    x = rng.standard_normal(nx).astype(np.float32)
    h = rng.standard_normal(nh).astype(np.float32)
    s_gpu = rng.standard_normal(nh - 1).astype(np.float32)
    s_cpu = s_gpu.copy()
    for _ in range(2):  # two consecutive blocks: state carry
        want = np.zeros(nx // decim, np.float32)
        orc.fir_decim(want, x, nx, h, nh, s_cpu, decim)
        got = sdr.convolveBlockFastFIR(x, h, s_gpu, decim) if decim > 1 else sdr.convolveBlockFIR(x, h, s_gpu)
        assert same(got, want)
        assert same(s_gpu, s_cpu)


def test_convolve_updown(sdr, orc):
    rng = np.random.default_rng(1)
    x = rng.standard_normal(500).astype(np.float32)
    h = rng.standard_normal(31).astype(np.float32)
    got = sdr.convolveFIR(x, h)
    # single-pass convolution == block convolution of x padded with nh-1 zeros, zero state
    xp = np.concatenate([x, np.zeros(30, np.float32)])
    want = np.zeros(xp.size, np.float32)
    orc.fir_block(want, xp, xp.size, h, h.size, np.zeros(30, np.float32))
    assert same(got, want)
    up = sdr.upsample(x, 3)
    assert up.size == 1500 and same(up[::3], x) and not up[1::3].any() and not up[2::3].any()
    assert same(sdr.downsample(x, 7), x[::7])


@pytest.mark.parametrize("U,D,tp,nx", [(3, 4, 17, 36), (147, 800, 13, 5600), (147, 800, 101, 5600),
                                        (441, 3200, 101, 22400), (7, 5, 9, 700)])
def test_fir_resample(sdr, orc, U, D, tp, nx):
    rng = np.random.default_rng(U + D)
    h = rng.standard_normal(tp * U).astype(np.float32)
    s_gpu = np.zeros(tp * U - 1, np.float32)
    s_cpu = s_gpu.copy()
    for _ in range(3):
        x = rng.standard_normal(nx).astype(np.float32)
        want = np.zeros(nx * U // D, np.float32)
        orc.fir_resample(want, x, nx, h, h.size, s_cpu, D, U)
        got = sdr.convolveBlockResampleFIR(x, h, s_gpu, D, U)
        assert same(got, want)
        assert same(s_gpu, s_cpu)


def test_fm_demod(sdr, orc):
    import ctypes as C
    rng = np.random.default_rng(2)
    n = 5120
    I = rng.standard_normal(n).astype(np.float32)
    Q = rng.standard_normal(n).astype(np.float32)
    I[100] = Q[100] = 0.0      # zero-denominator branch, filter.cpp:254
    I[0] = Q[0] = 0.0
    pi, pq = C.c_float(0.25), C.c_float(-0.5)
    want = np.zeros(n, np.float32)
    orc.fm_demod(want, I, Q, n, C.byref(pi), C.byref(pq))
    got, gpi, gpq = sdr.fmDemod(I, Q, 0.25, -0.5)
    assert same(got, want) and gpi == pi.value and gpq == pq.value
    assert got[100] == 0.0 and got[0] == 0.0


def test_allpass(sdr, orc):
    rng = np.random.default_rng(3)
    x = rng.standard_normal(5120).astype(np.float32)
    s_gpu = rng.standard_normal(75).astype(np.float32)
    s_cpu = s_gpu.copy()
    want = np.zeros_like(x)
    orc.allpass(x, x.size, s_cpu, 75, want)
    got = sdr.allPass(x, s_gpu)
    assert same(got, want) and same(s_gpu, s_cpu)


@pytest.mark.parametrize("freq,Fs,scale,adj,bw", [(19e3, 240e3, 2.0, 0.0, 0.01), (19e3, 288e3, 2.0, 0.0, 0.01),
                                                   (114e3, 240e3, 0.5, 1.178097, 0.002)])
def test_pll_bit_exact_over_many_blocks(sdr, orc, freq, Fs, scale, adj, bw):
    """The PLL is ill-conditioned: any 1-ulp difference in atan2f/sincosf/cosf or in the double
    phase expression diverges within a few blocks, so equality over 40 blocks (trigArg beyond
    1e5, i.e. deep in sincosf's large-argument reduction) is a strong check."""
    rng = np.random.default_rng(4)
    n = 5120
    st_gpu = np.array([0, 0, 1, 0, 1, 0], np.float32)
    st_cpu = st_gpu.copy()
    t0 = 0
    for blk in range(40):
        t = (t0 + np.arange(n)) / Fs
        x = (0.1 * np.sin(2 * np.pi * freq * t + 0.3) + 0.01 * rng.standard_normal(n)).astype(np.float32)
        if blk == 3:
            x[:50] = 0.0            # atan2(+-0, +-0) corner cases
        t0 += n
        want = np.zeros(n + 1, np.float32)
        orc.pll(x, n, want, st_cpu, freq, Fs, scale, adj, bw)
        got = sdr.fmPLL(x, st_gpu, freq, Fs, scale, adj, bw)
        assert same(got, want), f"block {blk}"
        assert same(st_gpu, st_cpu), f"block {blk}"


def test_filter_design_on_host_matches_oracle(sdr, orc):
    assert same(sdr.impulseResponseLPF(2.4e6, 1e5, 151), orc.lpf(2.4e6, 1e5, 151))
    assert same(sdr.bandPass(240e3, 18.5e3, 19.5e3, 151), orc.bpf(240e3, 18.5e3, 19.5e3, 151))
