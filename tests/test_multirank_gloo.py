"""World-size-2 run of the sharded path on CPU (gloo): each rank processes its contiguous slice of
batches, statistics are merged with one all-reduce, outputs are concatenated in rank order.  The
result must equal the single-process run bit for bit.  The per-rank engine here is the CPU oracle;
the GPU engines are sharded by exactly the same host logic (faqcs_b200/shard.py, bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from faqcs_b200 import shard, synth
from faqcs_b200.api import Options
from oracle_binding import OracleEngine

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, r1, r2, batch_records, okw, out_q):
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b1, b2 = shard.record_batches(r1, batch_records), shard.record_batches(r2, batch_records)
    lo, hi = shard.batch_slice(len(b1), world, rank)
    with OracleEngine(Options(**okw)) as eng:
        eng.autodetect(r1[b1[0][0]:b1[0][1]], r2[b2[0][0]:b2[0][1]])        # A1 runs on the FIRST batch on every rank
        streams = [b"", b"", b"", b""]
        for k in range(lo, hi):
            res = eng.process(r1[b1[k][0]:b1[k][1]], r2[b2[k][0]:b2[k][1]], k * batch_records, k == len(b1) - 1)
            for i in range(4):
                streams[i] += res.streams[i]
        merged = shard.allreduce_stats(eng.stats())
    gathered = [None] * world
    dist.all_gather_object(gathered, streams)
    if rank == 0:
        out_q.put((merged, [b"".join(g[i] for g in gathered) for i in range(4)]))
    dist.barrier()
    dist.destroy_process_group()


def test_batch_slices_cover_everything():
    for n in (0, 1, 5, 8, 17):
        for world in (1, 2, 4, 8):
            cuts = [shard.batch_slice(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            assert max(hi - lo for lo, hi in cuts) - min(hi - lo for lo, hi in cuts) <= 1


def test_world2_equals_single_process():
    w = synth.c2(6000)
    okw = dict(discard_output=True, quality=12)
    batch_records = 1000
    with OracleEngine(Options(**okw)) as eng:
        eng.autodetect(w.r1, w.r2)
        single = eng.process(w.r1, w.r2)
        single_stats = eng.stats()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, w.r1, w.r2, batch_records, okw, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, streams = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [bytes(s) for s in streams] == [bytes(s) for s in single.streams]
    assert not merged.diff(single_stats), merged.diff(single_stats)


def test_flatten_roundtrip_pads_rows():
    w = synth.c5(500)
    with OracleEngine(Options(quality=20)) as eng:
        eng.autodetect(w.r1)
        eng.process(w.r1)
        st = eng.stats()
    rows = shard.stats_row_counts(st) + 7
    back = shard.unflatten_stats(shard.flatten_stats(st, rows), rows, st)
    assert back.pre_quality_matrix.shape[0] == st.pre_quality_matrix.shape[0] + 7
    assert np.array_equal(back.pre_quality_matrix[:st.pre_quality_matrix.shape[0]], st.pre_quality_matrix)
    assert int(back.pre_quality_matrix[st.pre_quality_matrix.shape[0]:].sum()) == 0
    assert np.array_equal(back.filter_stats, st.filter_stats)
