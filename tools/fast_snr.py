"""Prints the SNR of every intermediate of SDR_VARIANT_FAST against the CPU oracle (per mode)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sdr_b200 as sdr, orclib
from sdr_b200 import siggen
orc = orclib.ORC()
def snr(ref, got):
    ref = ref.astype(np.float64); got = got.astype(np.float64)
    e = np.sum((ref - got) ** 2)
    return np.inf if e == 0 else 10 * np.log10(np.sum(ref ** 2) / e)
for mode in range(4):
    iq = siggen.make_batch(2, mode, 3, "stereo")
    with sdr.Pipeline(mode=mode, channels=1, batch=2, variant=sdr.VARIANT_FAST, max_bytes_per_channel=iq.shape[1]) as p:
        p.keep_taps(True)
        pcm = p.process_host(iq)
        got = {n: p.tap(n, 0) for n in ("i_filt", "q_filt", "demod", "audio_filt")}
    want_pcm, want = orc.run_chain(iq[0], mode, 1)
    d = np.abs(pcm[0].astype(np.int32) - want_pcm.astype(np.int32))
    print(mode, {n: round(snr(want[n], got[n]), 1) for n in got}, "pcm max diff", d.max(), "frac differing", round(float((d > 0).mean()), 4))
