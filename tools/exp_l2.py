"""Is the tensor-core front end bound by HBM latency?  Time k_rf_demod_tc for a batch that fits
in L2 (one wave of CTAs), back to back (input L2-resident) and with an L2 flush in between."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sdr_b200 as sdr

def run(batch, blocks, flush, mode=2, reps=12):
    info = sdr.mode_info(mode, 1)
    nbytes = blocks * info.block_bytes
    rng = np.random.default_rng(1)
    iq = torch.from_numpy(rng.integers(100, 156, (batch, nbytes), dtype=np.uint8)).cuda()
    p = sdr.Pipeline(mode=mode, channels=1, batch=batch, variant=sdr.VARIANT_FAST, max_bytes_per_channel=nbytes)
    pcm = torch.empty((batch, p.pcm_count(nbytes)), dtype=torch.int16, device="cuda")
    junk = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for r in range(reps + 3):
        if r == 3:
            p.profile(True); p.kernel_times(reset=True)
        if flush: junk.fill_(r & 255)
        p.process_device(iq.data_ptr(), nbytes, nbytes, pcm.data_ptr(), pcm.shape[1], st)
    torch.cuda.synchronize()
    kt = p.kernel_times()
    print(f"batch {batch} blocks {blocks} input {batch*nbytes/1e6:.0f} MB flush={flush}:",
          {k: round(v[0] / v[1], 4) for k, v in kt.items()})
    p.close()

for batch, blocks in ((98, 8), (98, 4), (49, 8)):
    for flush in (False, True):
        run(batch, blocks, flush)
