"""Phase timeline of one CTA of k_rf_demod_tc (needs a library built with SDR_TC_TRACE=1:
`SDR_TC_TRACE=1 python software-defined-radio_b200/build.py --force`).  Prints, per tile and worker
warp, the cycles spent in each phase."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sdr_b200 as sdr

mode, batch, blocks = 2, 1024, 16
info = sdr.mode_info(mode, 1)
nbytes = blocks * info.block_bytes
rng = np.random.default_rng(1)
iq = torch.from_numpy(rng.integers(100, 156, (batch, nbytes), dtype=np.uint8)).cuda()
p = sdr.Pipeline(mode=mode, channels=1, batch=batch, variant=sdr.VARIANT_FAST, max_bytes_per_channel=nbytes)
pcm = torch.empty((batch, p.pcm_count(nbytes)), dtype=torch.int16, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for r in range(3):
    p.process_device(iq.data_ptr(), nbytes, nbytes, pcm.data_ptr(), pcm.shape[1], st)
torch.cuda.synchronize()
out = np.zeros(9 * 16 * 12, np.int64)
rc = sdr.lib().sdr_debug_tc_trace(out.ctypes.data_as(C.c_void_p), out.size)
assert rc == 0
t = out.reshape(9, 16, 12)
nb = 296
tb, te = np.zeros(nb, np.int64), np.zeros(nb, np.int64)
assert sdr.lib().sdr_debug_tc_times(tb.ctypes.data_as(C.c_void_p), te.ctypes.data_as(C.c_void_p), nb) == 0
print("wall clock (ns): kernel span", te.max() - tb.min(), " CTA durations min/median/max", (te - tb).min(), int(np.median(te - tb)), (te - tb).max(),
      " start spread", tb.max() - tb.min(), " CTA 7:", te[7] - tb[7])
t0 = t[0, 15, 10]
names = ["top", "raw ok", "mma(it-1) ok", "transposed", "epi start", "epi mma ok", "I combined", "pre-barrier", "post-barrier", "epi end"]
print("CTA total:", t[0, 15, 11] - t0, " item starts (warp 0):", [int(v - t0) for v in t[0, 15, :10]])
print("third item: loop tops (warp 1):", [int(t[1, it, 0] - t0) for it in range(12)])
for it in (0, 1, 2, 9, 10, 11):
    print(f"--- third item, tile iteration {it} (cycles since CTA start; epilogue columns belong to tile {it - 1})")
    for w in (0, 1, 5):
        row = t[w, it, :10] - t0
        print(f"warp {w}: " + "  ".join(f"{n}={v}" for n, v in zip(names, row)))
    print(f"issuer: streams ready={t[8, it, 10] - t0} issued={t[8, it, 11] - t0}")
