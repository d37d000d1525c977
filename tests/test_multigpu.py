"""Multi-GPU product path on real devices (skipped unless >= 2 GPUs are visible): contiguous batch slices per device, the
library's own NCCL all-reduce (fq_allreduce_stats), outputs concatenated in device order.  Everything must equal the
one-GPU run and the CPU oracle bit for bit; the command-line driver with --devices must leave the reference's files."""
import os
import subprocess
import sys

import numpy as np
import pytest

import refcli
from faqcs_b200 import dist_stats, shard, synth
from faqcs_b200.api import Engine, Options
from oracle_binding import OracleEngine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


pytestmark = [pytest.mark.gpu, pytest.mark.skipif(_n_gpus() < 2, reason="needs at least two GPUs")]


@pytest.mark.parametrize("n_dev", [2, 4, 8])
def test_devices_in_one_process_equal_one_gpu_and_oracle(n_dev):
    if _n_gpus() < n_dev:
        pytest.skip(f"{n_dev} GPUs not visible")
    w = synth.c2(24000)
    okw = dict(discard_output=True, quality=12)
    batch_records = 1500
    b1, b2 = shard.record_batches(w.r1, batch_records), shard.record_batches(w.r2, batch_records)
    with Engine(Options(**okw)) as one, OracleEngine(Options(**okw)) as ora:
        one.autodetect(w.r1, w.r2)
        ora.autodetect(w.r1, w.r2)
        single = one.process(w.r1, w.r2)
        ref = ora.process(w.r1, w.r2)
        single_stats, ora_stats = one.stats(), ora.stats()
    assert [bytes(s) for s in single.streams] == [bytes(s) for s in ref.streams]
    engines = [Engine(Options(**okw), device=d) for d in range(n_dev)]
    try:
        streams = [b"", b"", b"", b""]
        for d, eng in enumerate(engines):
            eng.autodetect(w.r1[b1[0][0]:b1[0][1]], w.r2[b2[0][0]:b2[0][1]])      # A1 on the first batch, on every device
            lo, hi = shard.batch_slice(len(b1), n_dev, d)
            for k in range(lo, hi):
                res = eng.process(w.r1[b1[k][0]:b1[k][1]], w.r2[b2[k][0]:b2[k][1]], k * batch_records, k == len(b1) - 1)
                for i in range(4):
                    streams[i] += res.streams[i]
        ms = dist_stats.allreduce_local_engines(engines)
        assert ms >= 0.0
        for eng in engines:                                     # every device holds the merged block afterwards
            st = eng.stats()
            assert not st.diff(single_stats), st.diff(single_stats)
            assert not st.diff(ora_stats), st.diff(ora_stats)
    finally:
        for eng in engines:
            eng.close()
    assert [bytes(s) for s in streams] == [bytes(s) for s in single.streams]


def test_ranks_with_different_row_capacities_merge():
    """One device sees 600-base reads (its statistics block grows), the other 150-base reads: the collective must agree on the
    layout first (all-reducing blocks of different sizes hangs NCCL or mis-adds columns)."""
    w = synth.c2(3000)
    rng = np.random.default_rng(9)
    long_recs = []
    for i in range(200):
        L = int(rng.integers(400, 600))
        long_recs.append((f"@L{i}", "".join(rng.choice(list("ACGT"), size=L)), "I" * (L - 1) + "5"))
    long_r = np.frombuffer(synth.fastq_bytes(long_recs), dtype=np.uint8)
    okw = dict(input_quality_offset=33)
    with OracleEngine(Options(**okw)) as ora:
        ora.process(w.r1)
        ora.process(long_r)
        want = ora.stats()
    engines = [Engine(Options(**okw), device=0), Engine(Options(**okw), device=1)]
    try:
        engines[0].process(w.r1)
        engines[1].process(long_r)
        dist_stats.allreduce_local_engines(engines)
        for eng in engines:
            st = eng.stats()
            assert not st.diff(want), st.diff(want)
    finally:
        for eng in engines:
            eng.close()


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    w = synth.c2(12000)
    b1, b2 = shard.record_batches(w.r1, 1000), shard.record_batches(w.r2, 1000)
    lo, hi = shard.batch_slice(len(b1), world, rank)
    with Engine(Options(discard_output=True), device=rank) as eng:
        eng.autodetect(w.r1[b1[0][0]:b1[0][1]], w.r2[b2[0][0]:b2[0][1]])
        streams = [b"", b"", b"", b""]
        for k in range(lo, hi):
            res = eng.process(w.r1[b1[k][0]:b1[k][1]], w.r2[b2[k][0]:b2[k][1]], k * 1000, k == len(b1) - 1)
            for i in range(4):
                streams[i] += res.streams[i]
        dist_stats.allreduce_engine_stats(eng, dist, torch.device("cuda", rank))
        merged = eng.stats()
    gathered = [None] * world
    dist.all_gather_object(gathered, streams)
    if rank == 0:
        q.put((merged, [b"".join(g[i] for g in gathered) for i in range(4)]))
    dist.barrier()
    dist.destroy_process_group()


def test_one_process_per_gpu_over_nccl():
    """The torchrun shape (bench.py): each rank owns one Engine, the NCCL id travels over the process group."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    w = synth.c2(12000)
    with Engine(Options(discard_output=True)) as eng:
        eng.autodetect(w.r1, w.r2)
        single = eng.process(w.r1, w.r2)
        single_stats = eng.stats()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, streams = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert [bytes(x) for x in streams] == [bytes(x) for x in single.streams]
    assert not merged.diff(single_stats), merged.diff(single_stats)


@pytest.mark.skipif(not refcli.have_ref(), reason="reference binary not present")
def test_cli_with_two_devices_leaves_the_reference_files():
    from test_cli_dropin import assert_same_files, run_both
    w = synth.c2(60000)
    outs = run_both({"-1": ("r1.fq", w.r1), "-2": ("r2.fq", w.r2)}, ["--discard"], threads=4, extra_cli=["--batch_mb", "2", "--devices", "0,1"])
    assert_same_files(outs)
    w3 = synth.c3(40000)
    fa = "".join(f">{n}\n{s}\n" for n, s in w3.artifacts).encode()
    outs = run_both({"-1": ("r1.fq", w3.r1), "-2": ("r2.fq", w3.r2), "--artifactFile": ("primers.fa", fa)},
                    ["--adapter", "--polyA", "--rate", "0.2"], threads=3, extra_cli=["--devices", "0,1"])
    assert_same_files(outs)
