#!/usr/bin/env python
"""bench.py -- throughput of the FM receiver DSP hot path on B200.

Metric (BASELINE.json): aggregate input I/Q MS/s (10^6 complex samples, one I byte + one Q byte
each, consumed per second, summed over captures), device-timed, plus % of the HBM roofline.

Workload at N=1: BASELINE.json configs[1] -- mono custom-rate mode 2 (2.4 MS/s -> 44.1 kS/s through
the 147/800 polyphase resampler), 1024 batched captures on one B200, functional tap set
(rf 151 / audio 101 per phase).  One "step" = one pass of the whole path over one resident batch
([1024][blocks*112000] uint8, larger than L2).  N>1: every rank owns `--batch` captures on its own
GPU (captures are independent: no collective on the data path, weak scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W]          our CUDA path
      (also times, with short runs, the other BASELINE.json configs -> "other_configs"; --no-others skips them)
  python bench.py --impl reference [...]                       the reference's CPU path
  torchrun ... bench.py --gpus N ...                           one rank per GPU (driver-launched)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "aggregate input I/Q MS/s (device-timed)"
UNIT = "MS/s"
TAPS = dict(rf_taps=151, audio_taps=101, stereo_taps=151)


def algorithmic_bytes_per_sample(mode: int, channels: int) -> float:
    """2 B read (u8 I + u8 Q) + int16 PCM written per input complex sample."""
    fs = {0: 2.4e6, 1: 1.44e6, 2: 2.4e6, 3: 0.96e6}[mode]
    fa = {0: 48e3, 1: 48e3, 2: 44.1e3, 3: 44.1e3}[mode]
    return 2.0 + 2.0 * channels * fa / fs


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.rows = []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# --------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (oracle/_ref when it was compiled from
# /root/reference, else the oracle port), all host threads, bounded sample of the same workload
# --------------------------------------------------------------------------------------------
def _cpu_worker(args):
    kind, mode, channels, seed, n_blocks, reps, rds = args
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orclib
    from sdr_b200 import siggen
    lib = orclib.REF() if kind == "reference" else orclib.ORC()
    iq = siggen.make_capture(seed, mode, n_blocks, "rds" if rds else "stereo")
    rdo = orclib.RDS() if rds else None
    t0 = time.perf_counter()
    for _ in range(reps):
        _, taps = lib.run_chain(iq, mode, channels, TAPS["rf_taps"], TAPS["audio_taps"], TAPS["stereo_taps"],
                                keep_taps=rds)
        if rds:  # the RDS model has no C++ form: the oracle's C restatement of fmRDS.py follows
            fm = taps["demod"].astype(np.float64)
            rdo.run_chain(fm[:fm.size // 9600 * 9600], mode, 9600, keep=())
    return time.perf_counter() - t0, reps * iq.size // 2


def cpu_baseline(mode: int, channels: int, seconds: float = 12.0, cores: int | None = None,
                 rds: bool = False):
    """Times the CPU path on `cores` processes, one independent capture each (embarrassingly
    parallel, like the GPU batch).  Returns (MS/s, kind, cores, sample description)."""
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orclib
    import sdr_b200  # noqa: F401  (registers the package alias for the workers)
    kind = "reference" if orclib.REF() is not None else "port"
    cores = cores or os.cpu_count() or 1
    n_blocks = ({0: 15, 2: 12}[mode] if rds else 8)
    # calibrate one capture on one core, then size reps for ~`seconds` of work per core
    t, n = _cpu_worker((kind, mode, channels, 0, n_blocks, 1, rds))
    reps = max(1, int(seconds / max(t, 1e-3)))
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(kind, mode, channels, c, n_blocks, reps, rds) for c in range(cores)])
    wall = time.perf_counter() - t0
    total = sum(r[1] for r in res)
    sample = (f"{cores} captures x {n_blocks} reference blocks x {reps} passes, mode {mode}, "
              f"{'stereo' if channels == 2 else 'mono'}, taps 151/101/151, one process per core")
    if rds:
        sample += "; receiver chain as above + the oracle's C restatement of the Python RDS model"
        kind = "port"
    return total / wall / 1e6, kind, cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    per_step = max(1.0, min(20.0, 90.0 / (steps + args.warmup)))
    vals = []
    for i in range(args.warmup + steps):
        v, kind, cores, sample = cpu_baseline(args.mode, args.audio_channels, seconds=per_step, rds=args.rds)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    n_samples_step = None
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    if args.mixed:
        return {"workload": f"mode {args.mode} mixed: {args.batch // 2} mono + {args.batch // 2} stereo captures per "
                            f"GPU x {args.blocks} reference blocks ({args.blocks * BLOCK_BYTES[args.mode]} B each), "
                            f"{args.streams} pipeline handles on {args.streams} CUDA streams, taps 151/101/151",
                "mode": args.mode, "audio_channels": "1+2", "batch_per_gpu": args.batch,
                "blocks_per_capture": args.blocks, "streams": args.streams,
                "l2_policy": "input batch larger than L2 (no flush needed)",
                "variant": "mono: fast (+-1 LSB), stereo: exact (bit-identical)"}
    return {"workload": f"mode {args.mode} {'stereo' if args.audio_channels == 2 else 'mono'}, "
                        f"{args.batch} captures per GPU x {args.blocks} reference blocks "
                        f"({args.blocks * BLOCK_BYTES[args.mode]} B each), taps rf 151 / audio 101"
                        f"{' / stereo 151' if args.audio_channels == 2 else ''}"
                        f"{' + RDS chain (fmRDS.py model, double precision, 9600-sample blocks)' if args.rds else ''}"
                        f"{f', {args.streams} pipeline handles on {args.streams} CUDA streams' if args.streams > 1 else ''}",
            "mode": args.mode, "audio_channels": args.audio_channels, "batch_per_gpu": args.batch,
            "blocks_per_capture": args.blocks, "l2_policy": "input batch larger than L2 (no flush needed)",
            "variant": ("fast (tensor-core RF front end; PCM within +-1 LSB of the reference)"
                        if args.variant == "fast" and args.audio_channels == 1
                        else "mixed (bit-exact in front of the PLL, contracted multiply-adds elsewhere; PCM +-1 LSB)"
                        if args.variant == "mixed" else "exact (bit-identical to the reference)")}


BLOCK_BYTES = {0: 102400, 1: 61440, 2: 112000, 3: 134400}


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def make_host_batch(mode, batch, blocks, kind, out=None):
    """[batch, nbytes] uint8 on the host: a few distinct synthetic captures, tiled with a
    per-capture circular shift so that every row is a different valid capture."""
    from sdr_b200 import siggen
    distinct = min(batch, 8)
    base = [siggen.make_capture(c, mode, blocks, kind) for c in range(distinct)]
    n = base[0].size
    if out is None:
        out = np.empty((batch, n), np.uint8)
    for c in range(batch):
        rot = 2 * ((c // distinct) * 977 % (n // 2))
        src = base[c % distinct]
        if rot:
            out[c, :rot] = src[n - rot:]
            out[c, rot:] = src[:n - rot]
        else:
            out[c] = src
    return out


def make_device_batch(torch, mode, batch, blocks, kind, device):
    """A few distinct captures are built on the host; the batch is tiled from them on the device
    with the same per-row circular shift as make_host_batch (outside every timed region)."""
    from sdr_b200 import siggen
    distinct = min(batch, 8)
    base = torch.from_numpy(np.stack([siggen.make_capture(c, mode, blocks, kind) for c in range(distinct)])).to(device)
    n = base.shape[1]
    out = torch.empty((batch, n), dtype=torch.uint8, device=device)
    for g in range((batch + distinct - 1) // distinct):
        rot = 2 * (g * 977 % (n // 2))
        rows = min(distinct, batch - g * distinct)
        out[g * distinct:g * distinct + rows] = torch.roll(base[:rows], shifts=rot, dims=1) if rot else base[:rows]
    return out


def parity_rows(torch, orc_mod, groups, d_iq, nbytes, mode, rows_per_group=4):
    """After the timed region: reset every pipeline handle, run ONE call on the bench's own device
    batch and compare a few rows of its PCM with the CPU oracle fed with the same bytes (the
    oracle is test infrastructure; nothing of it runs inside a timed region).  Exact variant:
    bit-identical; fast / mixed: within +-1 LSB."""
    orc = orc_mod.ORC()
    worst, checked, ok = 0, 0, True
    per = groups[0][2].shape[0]
    for gi, (p, rds, rows, d_pcm, stream, vname) in enumerate(groups):
        if rds is not None:
            rds.discard()
        p.reset()
        d_pcm.zero_()
        torch.cuda.synchronize()
        p.process_device(rows.data_ptr(), rows.stride(0), nbytes, d_pcm.data_ptr(), d_pcm.stride(0), stream.cuda_stream)
        torch.cuda.synchronize()
        ch = p.cfg.channels
        pick = sorted({0, 1, per // 2, per - 1})[:rows_per_group]
        for r in pick:
            want, _ = orc.run_chain(rows[r].cpu().numpy(), mode, ch, TAPS["rf_taps"], TAPS["audio_taps"],
                                    TAPS["stereo_taps"], keep_taps=False)
            got = d_pcm[r].cpu().numpy()
            d = int(np.abs(got.astype(np.int32) - want.astype(np.int32)).max()) if got.size == want.size else 1 << 30
            worst = max(worst, d)
            ok = ok and d <= (0 if vname == "exact" else 1)
            checked += 1
    return {"parity_checked": bool(ok), "rows": checked, "max_abs_lsb": worst,
            "bar": "bit-identical (exact) / +-1 LSB (fast, mixed) vs the CPU oracle on the bench's own input"}


def pick_variant(sdr, args, mode, audio_channels):
    """FAST exists for mono (all modes); stereo needs the bit-exact chain in front of the PLL:
    EXACT (everything bit-identical) or MIXED (only what feeds the PLL is exact)."""
    if args.variant == "fast" and audio_channels == 1:
        return sdr.VARIANT_FAST, "fast"
    if args.variant == "mixed":
        return sdr.VARIANT_MIXED, "mixed"
    if args.variant == "exact_scalar":   # A/B: the FMUL + FADD form of the exact FIR kernels (same bits)
        return sdr.VARIANT_EXACT | sdr.VARIANT_SCALAR_FIR, "exact"
    return sdr.VARIANT_EXACT, "exact"


def fp32_ceiling_msps(mode, audio_channels, variant, sm_mhz):
    """What the CUDA-core (exact / mixed) chains are bounded by: FP32-pipe instructions per input
    sample of the FIR kernels alone (PLL, discriminator, staging, PCM packing counted as free) over
    148 SMs x 128 lanes x clock.  The reference's multiply-add is two instructions (separately
    rounded FMUL + FADD); a contracted one is one."""
    D = {0: 10, 1: 5, 2: 10, 3: 3}[mode]
    Da_over_U = {0: 5.0, 1: 6.0, 2: 800.0 / 147.0, 3: 3200.0 / 441.0}[mode]
    exact, fma = 2.0, 1.0
    rf = 2 * TAPS["rf_taps"] / D * (exact if variant != "fast" else 0.0)      # I and Q (fast: tensor cores)
    audio_cost = exact if variant == "exact" else fma
    audio = TAPS["audio_taps"] / (D * Da_over_U) * audio_cost * (2 if audio_channels == 2 else 1)
    bpf = 0.0
    if audio_channels == 2:
        bpf = TAPS["stereo_taps"] / D * (exact + (exact if variant == "exact" else fma))   # pilot (always exact) + 22-54 kHz band
    instr = rf + audio + bpf
    rate = 148 * 128 * (sm_mhz or 1965.0) * 1e6
    return {"fp32_instr_per_sample": instr, "ceiling_msps": rate / instr / 1e6 if instr else None}


def time_config(torch, sdr, args, mode, audio_channels, steps, warmup, dist, world, profile=True,
                force_exact=False, check_parity=False):
    """Device-resident timing of one configuration.  Returns dict with ms/step (max over ranks),
    per-kernel times and launches.

    With --streams S the batch is cut into S groups of captures, each owned by its own pipeline
    handle on its own CUDA stream (captures are independent, handles are independent): kernels
    that cannot fill the GPU on their own -- the sequential PLL, one lane per capture -- then run
    beside the other groups' kernels.  With --mixed the first half of the groups is mono (the
    fast variant unless --variant exact), the second half stereo: BASELINE.json configs[3]."""
    dev = torch.device("cuda", torch.cuda.current_device())
    kind = "rds" if args.rds else "stereo"
    d_iq = make_device_batch(torch, mode, args.batch, args.blocks, kind, dev)
    nbytes = d_iq.shape[1]
    S = max(1, args.streams)
    if args.batch % S or (args.mixed and S % 2):
        raise SystemExit("--batch must be a multiple of --streams (and --streams even with --mixed)")
    per = args.batch // S
    groups = []  # (pipeline, rds, rows, d_pcm, stream, variant name)
    main_stream = torch.cuda.current_stream()
    for i in range(S):
        ch = (1 if i < S // 2 else 2) if args.mixed else audio_channels
        variant, vname = pick_variant(sdr, args, mode, ch)
        if force_exact:
            variant, vname = sdr.VARIANT_EXACT, "exact"
        p = sdr.Pipeline(mode=mode, channels=ch, batch=per, device=dev.index,
                         max_bytes_per_channel=nbytes, variant=variant, **TAPS)
        rds = None
        if args.rds:
            # the RDS chain follows every process call on the same stream; its bit buffer is sized
            # for the whole run so that no host read-back falls inside the timed region
            if nbytes % 192000:
                raise SystemExit("--rds needs --blocks such that a capture is a multiple of 192000 B "
                                 "(mode 0: 15, 30, ...; mode 2: 12, 24, ...)")
            rds = sdr.Rds(p, block_if=9600, max_pending_blocks=(2 * steps + warmup) * (nbytes // 192000),   # timed pass + per-kernel pass
                          f32_fir=bool(getattr(args, "rds_f32", False)))
        n_pcm = p.pcm_count(nbytes)
        d_pcm = torch.zeros((per, n_pcm), dtype=torch.int16, device=dev)
        # alternating priorities: when two groups have throughput kernels ready, one goes first,
        # so the groups fall out of step and one group's PLL runs beside the other's FIR kernels
        stream = main_stream if S == 1 else torch.cuda.Stream(device=dev, priority=-(i % 2))
        groups.append((p, rds, d_iq[i * per:(i + 1) * per], d_pcm, stream, vname))

    def run(n_steps):
        """n_steps of every group; groups only meet at the start and at the end (no barrier per
        step: a group's next call may start while another group is still in its PLL)."""
        if S > 1:
            fork = torch.cuda.Event()
            fork.record(main_stream)
        for (p, _, rows, d_pcm, stream, _) in groups:
            if S > 1:
                stream.wait_event(fork)
            for _ in range(n_steps):
                p.process_device(rows.data_ptr(), rows.stride(0), nbytes, d_pcm.data_ptr(), d_pcm.stride(0),
                                 stream.cuda_stream)
            if S > 1:
                j = torch.cuda.Event()
                j.record(stream)
                main_stream.wait_event(j)

    run(warmup)
    torch.cuda.synchronize()
    for g in groups:
        g[0].launch_count(reset=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    # the timed region holds the product's own launches and nothing else
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(steps)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    launches = sum(g[0].launch_count() for g in groups)
    # per-kernel durations: a second pass of the same `steps` steps on the same inputs with a CUDA-event
    # pair around every kernel on its launching stream (kept out of the timed region above: the event
    # records sit between dependent kernels and cost ~1 us each)
    ktimes = {}
    ms_profiled = None
    if profile:
        for g in groups:
            g[0].profile(True)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        run(steps)
        p1.record()
        torch.cuda.synchronize()
        ms_profiled = p0.elapsed_time(p1)
        for g in groups:  # summed over the groups (with S > 1 they overlap in time)
            for k, (t, n) in g[0].kernel_times(reset=True).items():
                a = ktimes.get(k, (0.0, 0))
                ktimes[k] = (a[0] + t, a[1] + n)
    for g in groups:
        g[0].profile(False)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    samples_per_step = args.batch * (nbytes // 2)
    checksum = int(sum(int(g[3][:, :64].to(torch.int64).abs().sum().item()) for g in groups))
    rds_info = None
    if groups[0][1] is not None:
        rd = groups[0][1].read(0)
        rds_info = {"blocks": int(rd["bit_counts"].size), "bits_capture0": int(rd["cdr_bits"].size),
                    "offsets_capture0_tail": rd["offsets"][-16:]}
    parity = None
    if check_parity:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import orclib
        parity = parity_rows(torch, orclib, groups, d_iq, nbytes, mode)
    for g in groups:
        if g[1] is not None:
            g[1].close()
        g[0].close()
    n_pcm_total = sum(int(g[3].shape[1]) * per for g in groups)
    vname = "+".join(sorted({g[5] for g in groups}))
    del d_iq, groups
    torch.cuda.empty_cache()
    return {"ms_per_step": ms / steps, "samples_per_step": samples_per_step, "launches": launches,
            "kernels": ktimes, "nbytes": nbytes, "n_pcm": n_pcm_total // args.batch, "checksum": checksum,
            "variant": vname, "rds": rds_info, "pcm_values_per_step": n_pcm_total, "parity": parity,
            "ms_per_step_with_kernel_events": (ms_profiled / steps) if ms_profiled else None}


def copy_ceiling(torch, devices, bytes_per_device, reps=3):
    """Plain pinned host -> device copies, one cudaMemcpyAsync per device and repetition, all
    devices concurrently: what the box's host memory system and PCIe links deliver with no
    pipeline behind them.  Returns aggregate GB/s (the best of `reps` rounds)."""
    bufs = []
    for d in devices:
        h = torch.empty(bytes_per_device, dtype=torch.uint8).pin_memory()
        h.fill_(1)
        with torch.cuda.device(d):
            bufs.append((h, torch.empty(bytes_per_device, dtype=torch.uint8, device=f"cuda:{d}"),
                         torch.cuda.Stream(device=d)))
    best = 0.0
    for _ in range(reps + 1):
        for d in devices:
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        for (h, g, st) in bufs:
            with torch.cuda.stream(st):
                g.copy_(h, non_blocking=True)
        for d in devices:
            torch.cuda.synchronize(d)
        dt = time.perf_counter() - t0
        best = max(best, len(devices) * bytes_per_device / dt / 1e9)
    del bufs
    return best


def time_e2e(torch, sdr, args, steps, world, rank, barrier):
    """Same metric through the public host-buffer API, host buffers in, host PCM out, every copy
    inside the timed region.  One GPU: sdr_pipeline_process_host.  N GPUs: ONE process (rank 0)
    hands the whole batch of N x `batch` captures to sdr_multi_process_host, which spreads it over
    the N devices with one host thread each and gathers the PCM into one host array -- the final
    host gather is therefore inside the timed region; the other ranks only wait."""
    blocks = min(args.blocks, args.e2e_blocks)
    nbytes = blocks * BLOCK_BYTES[args.mode]
    if rank != 0:
        barrier()
        return None
    total = args.batch * world
    os.sched_setaffinity(0, range(os.cpu_count()))   # worker threads inherit: let them use every core
    h_iq = torch.empty((total, nbytes), dtype=torch.uint8).pin_memory()
    make_host_batch(args.mode, args.batch, blocks, "stereo", out=h_iq.numpy()[:args.batch])
    for r in range(1, world):   # the other devices' ranges: the same captures, rotated by one row
        h_iq[r * args.batch:(r + 1) * args.batch] = torch.roll(h_iq[:args.batch], shifts=r, dims=0)
    variant = pick_variant(sdr, args, args.mode, args.audio_channels)[0]
    if world == 1:
        p = sdr.Pipeline(mode=args.mode, channels=args.audio_channels, batch=total, device=torch.cuda.current_device(),
                         max_bytes_per_channel=nbytes, variant=variant, **TAPS)
        api = "sdr_pipeline_process_host (pinned host buffers, 3 streams, double-buffered slices)"
    else:
        p = sdr.MultiPipeline(mode=args.mode, channels=args.audio_channels, batch=total, devices=world,
                              max_bytes_per_channel=nbytes, variant=variant, **TAPS)
        api = (f"sdr_multi_process_host: one process, {world} devices, one host thread per device, contiguous "
               "capture ranges, PCM gathered into one pinned host array")
    n_pcm = p.pcm_count(nbytes)
    h_pcm = torch.empty((total, n_pcm), dtype=torch.int16).pin_memory()

    def step():
        p.process_host_ptr(h_iq.data_ptr(), h_iq.stride(0), nbytes, h_pcm.data_ptr(), h_pcm.stride(0))

    step()  # warm-up: allocates the staging buffers
    step()
    step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()          # returns when the whole batch's PCM is in h_pcm
    sec = time.perf_counter() - t0
    p.close()
    ceiling = copy_ceiling(torch, list(range(world)), args.batch * nbytes)
    del h_iq, h_pcm
    barrier()
    samples = total * (nbytes // 2) * steps
    h2d = int(total * nbytes)
    return {"value": samples / sec / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": int(total * n_pcm * 2), "blocks_per_capture": blocks,
            "ms_per_step": sec / steps * 1e3, "api": api,
            "h2d_gbs": h2d * steps / sec / 1e9,
            "copy_ceiling_gbs": ceiling,
            "frac_of_copy_ceiling": h2d * steps / sec / 1e9 / ceiling if ceiling else None,
            "copy_ceiling_how": f"{world} concurrent pinned cudaMemcpyAsync H2D of {args.batch * nbytes} B each, best of 3 "
                                "(tools/h2d_ceiling.cu is the stand-alone probe)",
            "host_gather_in_timed_region": True}


def bind_to_gpu_numa_node(index: int):
    """Pin this rank to the CPUs next to its GPU (NVML's affinity mask) before any pinned host
    memory is allocated, so that the end-to-end leg's H2D copies read local memory: with one
    rank per GPU the ranks otherwise share whichever NUMA node they happened to start on."""
    before = os.sched_getaffinity(0)
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        cpus &= before
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass  # no NVML / no permission: keep the inherited affinity
    return before  # the caller restores it before the CPU-baseline leg forks its workers


def time_channelizer(torch, sdr, peak, M=16, T=8, W=64, blocks=16, steps=5, warmup=3):
    """W wideband captures (M x 2.4 MS/s, uint8 I/Q) -> W*M receiver inputs in HBM
    (mode-0 captures of `blocks` reference blocks each).  Device-timed with CUDA events; algorithmic
    bytes: 2 read + 2 written per wideband I/Q pair."""
    dev = torch.device("cuda", torch.cuda.current_device())
    nbytes_ch = blocks * 102400
    nbytes_wide = nbytes_ch * M
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    d_wide = torch.randint(0, 256, (W, nbytes_wide), dtype=torch.uint8, device=dev, generator=g)
    d_out = torch.empty((W * M, nbytes_ch), dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    with sdr.Channelizer(M, T, n_wide=W, device=dev.index, gain=1.0) as ch:
        for _ in range(warmup):
            ch.process_device(d_wide.data_ptr(), nbytes_wide, nbytes_wide, d_out.data_ptr(), nbytes_ch, s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            ch.process_device(d_wide.data_ptr(), nbytes_wide, nbytes_wide, d_out.data_ptr(), nbytes_ch, s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    # parity on the bench's own input, outside the timed region: a fresh bank (empty history) over the first
    # 2 MiB of wideband row 0 against the double-precision CPU oracle (oracle/aux_oracle.c; test infrastructure)
    parity = None
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import auxlib
        n_chk = 2 * M * 65536
        wide0 = d_wide[0, :n_chk].cpu().numpy()
        with sdr.Channelizer(M, T, n_wide=1, device=dev.index, gain=1.0) as ch:
            h = ch.prototype()
            got = ch.process_host(wide0[None, :])
        want = auxlib.channelize(wide0, M, h, 1.0)
        dd = np.abs(got.astype(np.int32) - want.astype(np.int32))
        parity = {"parity_checked": bool(int(dd.max()) <= 1), "max_abs_lsb": int(dd.max()),
                  "mismatch_frac": float(np.mean(dd != 0)), "values": int(dd.size),
                  "bar": "<= 1 LSB of the 8-bit output vs the double-precision oracle (single vs double at rounding boundaries)"}
    except Exception as e:   # the check must not take the measurement down
        parity = {"parity_checked": False, "error": str(e)[:200]}
    pairs = W * nbytes_wide // 2
    gbs = pairs * 4 / (ms * 1e-3) / 1e9
    return {"what": f"SURVEY 8f3 channeliser: {W} wideband captures ({M} x 2.4 MS/s) -> {W * M} receiver inputs "
                    f"({blocks} reference blocks each) in HBM, {T} taps per branch",
            "value": pairs / (ms * 1e-3) / 1e6, "unit": "wideband input I/Q MS/s (device-timed)", "ms_per_step": ms,
            "steps": steps, "warmup": warmup, "algorithmic_bytes_per_pair": 4, "achieved_gbs": gbs,
            "frac_of_hbm_peak": gbs / peak, "gpu_launches": 2 * steps, "parity": parity}


def with_args(args, **kw):
    d = dict(vars(args))
    d.update(kw)
    return argparse.Namespace(**d)


def other_config_list(args, world):
    """The BASELINE.json configs that are not the headline one, as short runs (3 warm-up + 3 timed
    steps).  Per GPU; under torchrun only the sharded mixed config (configs[3]) and the stereo
    shapes are repeated on every rank."""
    base = dict(streams=1, mixed=False, rds=False, variant="fast", batch=1024, blocks=16)
    lst = []
    if world == 1:
        lst += [("mono_mode0", "configs[1]: mono, mode 0 (48 kS/s, plain decimating audio filter)",
                 dict(base, mode=0, audio_channels=1)),
                ("mono_mode1", "configs[1]: mono, mode 1 (1.44 MS/s, rf_decim 5)", dict(base, mode=1, audio_channels=1)),
                ("mono_mode3", "configs[1]: mono custom-rate mode 3 (960 kS/s -> 44.1 kS/s, 441/3200)",
                 dict(base, mode=3, audio_channels=1)),
                ("mono_mode2_exact", "headline shape on the bit-exact CUDA-core path",
                 dict(base, mode=2, audio_channels=1, variant="exact")),
                ("stereo_mode0_1024", "configs[2]: stereo, 1024 captures (the PLL's 32 warps cannot fill the GPU)",
                 dict(base, mode=0, audio_channels=2, variant="exact"))]
    lst += [("stereo_mode0_fill", "configs[2] shaped to fill the GPU: 32768 captures x 2 blocks (two PLL warps per "
                                  "scheduler), bit-exact",
             dict(base, mode=0, audio_channels=2, variant="exact", batch=32768, blocks=2)),
            ("stereo_mode0_fill_mixed", "same shape, SDR_VARIANT_MIXED (PLL chain exact, other filters contracted, PCM +-1 LSB)",
             dict(base, mode=0, audio_channels=2, variant="mixed", batch=32768, blocks=2)),
            ("mixed_8192", f"configs[3]: 8192 captures (half mono fast, half stereo exact) over {world} GPU(s), "
                           f"{8192 // world} per GPU, two pipeline handles on two streams",
             dict(base, mode=0, audio_channels=1, mixed=True, streams=2, batch=8192 // world, blocks=4))]
    if world == 1:
        lst += [("stereo_rds_mode0", "configs[4]: stereo + RDS chain (BPF, squaring PLL, RRC, clock recovery), 1024 captures x 15 blocks",
                 dict(base, mode=0, audio_channels=2, variant="exact", rds=True, blocks=15)),
                ("stereo_rds_fill", "configs[4] shaped to fill the GPU: 8192 captures x 15 blocks (the stereo PLL's 27 ms are "
                                    "flat up to 16 k captures; the RDS chain's FP64 filters scale with the batch)",
                 dict(base, mode=0, audio_channels=2, variant="exact", rds=True, blocks=15, batch=8192))]
    return lst


def run_ours(args):
    import torch
    import torch.distributed as dist
    import sdr_b200 as sdr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or sdr.device_count() < 1:
        raise SystemExit("bench.py needs a B200 (sm_100) GPU: the product has no CPU fallback")
    torch.cuda.set_device(local)
    affinity_before = bind_to_gpu_numa_node(local)
    cpu_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")   # host-side waits that keep the GPUs idle

    def cpu_barrier():
        if world > 1:
            dist.barrier(group=cpu_group)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    main = time_config(torch, sdr, args, args.mode, args.audio_channels, args.steps, args.warmup, dist, world,
                       check_parity=(rank == 0 and not args.no_parity))
    clocks = None
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=3)
        clocks = sampler.summary()
    peak, peak_src = measured_peak()
    sm_mhz = (clocks or {}).get("sm_mhz")

    others = {}
    if not args.no_others:
        for name, what, kw in other_config_list(args, world):
            a = with_args(args, **kw)
            try:
                r = time_config(torch, sdr, a, a.mode, a.audio_channels, 3, 3, dist, world, profile=True,
                                check_parity=(rank == 0 and not args.no_parity))
            except Exception as e:   # an optional line must not take the headline down with it
                others[name] = {"what": what, "error": str(e)[:300]}
                if world > 1:
                    raise
                continue
            msps = r["samples_per_step"] * world / (r["ms_per_step"] * 1e-3) / 1e6
            bps = algorithmic_bytes_per_sample(a.mode, a.audio_channels)
            if a.mixed:
                bps = 0.5 * (algorithmic_bytes_per_sample(a.mode, 1) + algorithmic_bytes_per_sample(a.mode, 2))
            kms = {k: v[0] / 3 for k, v in r["kernels"].items()}
            dom = max(kms, key=kms.get) if kms else None
            o = {"what": what, "value": msps, "unit": UNIT, "n_gpus": world, "variant": r["variant"],
                 "ms_per_step": r["ms_per_step"], "steps": 3, "warmup": 3,
                 "workload": workload_config(a)["workload"],
                 "whole_step_frac": msps * 1e6 * bps / 1e9 / peak / world,
                 "dominant_kernel": dom, "kernel_ms_per_step": kms, "gpu_launches": r["launches"],
                 "parity": r["parity"]}
            if a.audio_channels == 2 and not a.mixed:
                c = fp32_ceiling_msps(a.mode, 2, r["variant"], sm_mhz)
                o["fp32_pipe_ceiling"] = dict(c, frac_of_ceiling=msps / world / c["ceiling_msps"],
                                              note="FIR kernels' FP32-pipe instructions only (the reference's "
                                                   "multiply-add = FMUL + FADD), 148 SMs x 128 lanes x SM clock; "
                                                   "the bit-exact chain cannot reach the HBM roofline")
            if r.get("rds"):
                o["rds"] = r["rds"]
            others[name] = o

    if not args.no_others and world == 1:
        try:
            others["channelizer_16"] = time_channelizer(torch, sdr, peak)
        except Exception as e:
            others["channelizer_16"] = {"error": str(e)[:300]}

    e2e = None
    if not (args.rds or args.mixed or args.streams > 1):
        os.sched_setaffinity(0, affinity_before)
        if world == 1:
            bind_to_gpu_numa_node(local)
        e2e = time_e2e(torch, sdr, args, max(2, min(args.steps, 5)), world, rank, cpu_barrier)
    os.sched_setaffinity(0, affinity_before)

    if rank == 0:
        total_samples = main["samples_per_step"] * world
        value = total_samples / (main["ms_per_step"] * 1e-3) / 1e6
        bps = algorithmic_bytes_per_sample(args.mode, args.audio_channels)
        if args.mixed:  # 2 B in + the PCM both halves write
            bps = 0.5 * (algorithmic_bytes_per_sample(args.mode, 1) + algorithmic_bytes_per_sample(args.mode, 2))
        # dominant kernel: the one with the largest total time in the timed region
        kt = main["kernels"]
        dom = max(kt, key=lambda k: kt[k][0]) if kt else None
        step_kernel_ms = sum(v[0] for v in kt.values()) / args.steps if kt else None
        roofline = None
        if dom:
            dom_ms = kt[dom][0] / kt[dom][1]
            alg_bytes = main["samples_per_step"] * bps  # per launch: one launch covers the whole batch
            achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
            traffic, traffic_src = None, None
            try:  # DRAM bytes per launch of this kernel from the committed ncu --set full capture
                tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                w = tj["workload"]
                if (w["mode"], w["audio_channels"], w["batch_per_gpu"], w["blocks_per_capture"]) == \
                        (args.mode, args.audio_channels, args.batch, args.blocks) and dom in tj:
                    traffic = tj[dom]["dram_bytes_per_launch"]
                    traffic_src = ("not measured in this run: " + tj[dom]["source"])
            except Exception:
                pass
            roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                        "traffic_gbs": (traffic / (dom_ms * 1e-3) / 1e9) if traffic else None,
                        "traffic_frac_of_peak": (traffic / (dom_ms * 1e-3) / 1e9 / peak) if traffic else None,
                        "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": dom_ms,
                        "kernel_share_of_step": kt[dom][0] / args.steps / step_kernel_ms,
                        "whole_step_frac": value * 1e6 * bps / 1e9 / peak / world,
                        "kernel_ms_per_step": {k: v[0] / args.steps for k, v in kt.items()},
                        "kernel_timing": "CUDA-event pair around every kernel on its launching stream, second pass of the "
                                         "same K steps on the same inputs right after the timed region (the event records "
                                         "are kept out of the timed region)",
                        "ms_per_step_with_kernel_events": main.get("ms_per_step_with_kernel_events")}
        cb = None
        if roofline and args.streams > 1:
            roofline["note"] = ("per-kernel times are summed over the pipeline handles, which overlap in "
                                "time: shares of the step do not apply; whole_step_frac is the measured one")
        if world == 1 and not args.no_cpu_baseline and not args.mixed:
            v, kind, cores, sample = cpu_baseline(args.mode, args.audio_channels, seconds=10.0, rds=args.rds)
            cb = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args), "roofline": roofline, "cpu_baseline": cb, "e2e": e2e,
            "gpu_launches": main["launches"], "clocks": clocks, "pcm_checksum": main["checksum"],
        }
        if main.get("parity"):
            line["parity_checked"] = main["parity"]["parity_checked"]
            line["parity"] = main["parity"]
        if main.get("rds"):
            line["rds"] = main["rds"]
        if others:
            line["other_configs"] = others
            for k in ("stereo_mode0_fill", "stereo_mode0_1024"):   # stable top-level key for the stereo result
                if k in others and "value" in others[k]:
                    line["stereo"] = dict(others[k], config_key=k)
                    break
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", type=int, default=2)
    ap.add_argument("--audio-channels", type=int, default=1)
    ap.add_argument("--batch", type=int, default=1024, help="captures per GPU")
    ap.add_argument("--blocks", type=int, default=16, help="reference blocks per capture per step")
    ap.add_argument("--e2e-blocks", type=int, default=8)
    ap.add_argument("--no-others", action="store_true",
                    help="skip the short runs of the other BASELINE.json configs (other_configs)")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-run PCM check against the CPU oracle")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=1,
                    help="pipeline handles per GPU, each with batch/streams captures on its own CUDA stream")
    ap.add_argument("--mixed", action="store_true",
                    help="half of the handles mono, half stereo (BASELINE configs[3]); needs --streams >= 2")
    ap.add_argument("--rds-f32", action="store_true",
                    help="with --rds: the RDS chain's three FIR stages in single precision (experiment; DESIGN.md section 6)")
    ap.add_argument("--rds", action="store_true",
                    help="also run the RDS chain (modes 0/2) behind every step; not the default workload")
    ap.add_argument("--only-channelizer", action="store_true", help="time the channeliser alone and exit")
    ap.add_argument("--variant", default="fast", choices=["fast", "exact", "mixed", "exact_scalar"],
                    help="fast: tensor-core RF front end (mono, +-1 LSB PCM); exact: bit-identical CUDA-core path; "
                         "mixed: exact in front of the PLL, contracted multiply-adds elsewhere (+-1 LSB PCM)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.mixed and args.streams < 2:
        args.streams = 2
    if args.only_channelizer:
        import torch
        import sdr_b200 as sdr
        print(json.dumps(time_channelizer(torch, sdr, measured_peak()[0])))
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
