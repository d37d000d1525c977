"""Multi-GPU sharding of the hot path.

Reads (pairs) are independent and every statistic is a sum (matrix.h:111-142, trim.cpp:47-65,
120-154), so the path shards with no data-path collective: rank g of N owns a contiguous slice
of the input batches, keeps R1 and R2 of a pair together, and emits its own slice of the four
output streams; the host concatenates the slices in rank order, which is input order.  The only
exchange is one all-reduce (SUM) of the flat integer statistics plus a MAX of the row counters
at the end of the run -- `torch.distributed` over NCCL/NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from .api import NUM_BASE, NUM_QUAL, Stats


def batch_slice(n_batches: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of batch indices owned by `rank` (sizes differ by at most one)."""
    q, r = divmod(n_batches, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def record_batches(buf: np.ndarray, batch_records: int) -> List[Tuple[int, int]]:
    """Byte ranges of consecutive batches of `batch_records` whole FASTQ records (host-side batching:
    the reader only counts line ends, as FaQCs.cpp:240-251 does through next_read)."""
    nl = np.flatnonzero(np.asarray(buf) == 10)
    assert nl.size % 4 == 0, "FASTQ record count is not whole"
    ends = nl[3::4] + 1
    cuts = [0] + [int(ends[i]) for i in range(batch_records - 1, ends.size, batch_records)]
    if ends.size and cuts[-1] != int(ends[-1]):
        cuts.append(int(ends[-1]))
    return list(zip(cuts[:-1], cuts[1:]))


_POS_FIELDS = ("pre_quality_matrix", "post_quality_matrix", "pre_base_matrix", "post_base_matrix")
_LEN_FIELDS = ("pre_length_hist", "post_length_hist")
_FIXED_FIELDS = ("filter_stats", "adapter_reads", "adapter_bases", "pre_read_quality_hist", "pre_base_quality_hist",
                 "post_read_quality_hist", "post_base_quality_hist", "pre_composition", "post_composition")


def stats_row_counts(st: Stats) -> np.ndarray:
    return np.array([st.pre_quality_matrix.shape[0], st.post_quality_matrix.shape[0],
                     st.pre_length_hist.size, st.post_length_hist.size], dtype=np.int64)


def flatten_stats(st: Stats, rows: Sequence[int]) -> np.ndarray:
    """Stats -> one int64 vector whose layout depends only on the agreed row counts."""
    pre_rows, post_rows, pre_len, post_len = (int(x) for x in rows)
    parts = [np.asarray(getattr(st, f), dtype=np.int64).reshape(-1) for f in _FIXED_FIELDS]
    for f, nrow in zip(_POS_FIELDS, (pre_rows, post_rows, pre_rows, post_rows)):
        a = np.asarray(getattr(st, f), dtype=np.int64)
        pad = np.zeros((nrow, a.shape[1]), dtype=np.int64)
        pad[:a.shape[0]] = a
        parts.append(pad.reshape(-1))
    for f, n in zip(_LEN_FIELDS, (pre_len, post_len)):
        a = np.asarray(getattr(st, f), dtype=np.int64)
        pad = np.zeros(n, dtype=np.int64)
        pad[:a.size] = a
        parts.append(pad)
    return np.concatenate(parts)


def unflatten_stats(vec: np.ndarray, rows: Sequence[int], like: Stats) -> Stats:
    pre_rows, post_rows, pre_len, post_len = (int(x) for x in rows)
    out, o = {}, 0
    for f in _FIXED_FIELDS:
        shape = getattr(like, f).shape
        n = int(np.prod(shape))
        out[f] = vec[o:o + n].reshape(shape).astype(np.uint64)
        o += n
    for f, nrow, ncol in zip(_POS_FIELDS, (pre_rows, post_rows, pre_rows, post_rows), (NUM_QUAL, NUM_QUAL, NUM_BASE, NUM_BASE)):
        out[f] = vec[o:o + nrow * ncol].reshape(nrow, ncol).astype(np.uint64)
        o += nrow * ncol
    for f, n in zip(_LEN_FIELDS, (pre_len, post_len)):
        out[f] = vec[o:o + n].astype(np.uint64)
        o += n
    return Stats(**out)


def allreduce_stats(st: Stats) -> Stats:
    """Merge the per-rank statistics: MAX of the row counters, then SUM of the flat vector."""
    import torch
    import torch.distributed as dist
    rows = torch.from_numpy(stats_row_counts(st))
    dist.all_reduce(rows, op=dist.ReduceOp.MAX)
    vec = torch.from_numpy(flatten_stats(st, rows.tolist()))
    dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    return unflatten_stats(vec.numpy(), rows.tolist(), st)
