"""CPU tests (gloo, world_size 2) of the N>1 host logic: capture sharding and the final PCM gather.
The per-rank compute is stood in for by the CPU oracle here (tests may use it; the product never
does) -- on a GPU box each rank runs its own sdr_pipeline over the same shard."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import orclib
from sdr_b200 import sharding, siggen


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 8, 1024, 8191):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def _worker(rank, world, port, n_captures, result_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = orclib.ORC()
        b, e = sharding.shard_range(n_captures, world, rank)
        rows = []
        for c in range(b, e):
            iq = siggen.make_capture(c, 0, 1, "stereo")
            pcm, _ = orc.run_chain(iq, 0, 2, 13, 13, 13, keep_taps=False)
            rows.append(pcm)
        local = np.stack(rows) if rows else np.zeros((0, 2048), np.int16)
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == world
        full = sharding.gather_pcm(local, n_captures, dist)
        if rank == 0:
            np.save(result_path, full)
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


def test_two_ranks_gather_equals_single_process(tmp_path):
    n = 5  # odd on purpose: shards of 3 and 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "pcm.npy")
    mp.spawn(_worker, args=(2, port, n, out), nprocs=2, join=True)
    got = np.load(out)
    orc = orclib.ORC()
    want = np.stack([orc.run_chain(siggen.make_capture(c, 0, 1, "stereo"), 0, 2, 13, 13, 13, keep_taps=False)[0]
                     for c in range(n)])
    assert got.shape == want.shape and np.array_equal(got, want)


def _rds_reads(orc, captures):
    """What sdr_b200.Rds.read(c) returns, produced by the CPU oracles (the test's stand-in for
    the per-rank GPU chain)."""
    R = orclib.RDS()
    out = []
    for c in captures:
        iq = siggen.make_capture(c, 0, 15, "rds")
        _, taps = orc.run_chain(iq, 0, 1, 13, 13, 13)
        r = R.run_chain(taps["demod"].astype(np.float64), 0, 9600, keep=())
        out.append(dict(cdr_bits=np.concatenate(r["cdr_bits"]), diff_bits=np.concatenate(r["diff_bits"]),
                        bit_counts=np.array([b.size for b in r["cdr_bits"]], np.int32), offsets=r["offsets"]))
    return out


def _rds_worker(rank, world, port, n_captures, result_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = sharding.shard_range(n_captures, world, rank)
        full = sharding.gather_rds(_rds_reads(orclib.ORC(), range(b, e)), n_captures, dist)
        if rank == 0:
            np.save(result_path, np.array(full, dtype=object), allow_pickle=True)
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


def test_two_ranks_gather_rds_bits(tmp_path):
    n = 3  # shards of 2 and 1
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "rds.npy")
    mp.spawn(_rds_worker, args=(2, port, n, out), nprocs=2, join=True)
    got = list(np.load(out, allow_pickle=True))
    want = _rds_reads(orclib.ORC(), range(n))
    assert len(got) == n
    for g, w in zip(got, want):
        assert g["offsets"] == w["offsets"]
        for k in ("cdr_bits", "diff_bits", "bit_counts"):
            assert np.array_equal(g[k], w[k]), k
