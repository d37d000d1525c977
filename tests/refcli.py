"""Run the UNMODIFIED reference binary (oracle/_ref/FaQCs) and parse what it writes.

Used to pin the oracle (tests marked ``ref``) and to (re)generate the committed
fixtures under tests/golden/ (tests/golden/make_golden.py).
"""
import os
import shutil
import signal
import subprocess
import tempfile
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from faqcs_b200.api import (MODE_BWA, MODE_BWA_PLUS, MODE_HARD, NUM_COMPOSITION, NUM_COMPOSITION_BIN, NUM_QUAL,
                            OFFSET_AUTO, Options, BUILTIN_ADAPTERS, POLYA_ADAPTER)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "FaQCs")


def have_ref() -> bool:
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def flags_for(opt: Options, polyA: bool = False, artifact_file: Optional[str] = None, adapter: Optional[bool] = None) -> List[str]:
    """Options -> reference command line (options.cpp:138-178)."""
    f: List[str] = ["--mode", {MODE_HARD: "HARD", MODE_BWA: "BWA", MODE_BWA_PLUS: "BWA_plus"}[opt.mode],
                    "-q", str(opt.quality), "--min_L", str(opt.min_read_length), "-n", str(opt.max_num_poly_N),
                    "--avg_q", repr(float(opt.average_quality)), "--lc", repr(float(opt.low_complexity_cutoff_ratio)),
                    "--rate", repr(float(opt.adapter_mismatch_rate)), "--out_ascii", str(opt.output_quality_offset)]
    if opt.trim_5:
        f += ["--5end", str(opt.trim_5)]
    if opt.trim_3:
        f += ["--3end", str(opt.trim_3)]
    if opt.input_quality_offset != OFFSET_AUTO:
        f += ["--ascii", str(opt.input_quality_offset)]
    if opt.replace_to_N_q:
        f += ["--replace_to_N_q", str(opt.replace_to_N_q)]
    if opt.qc_only:
        f += ["--qc_only"]
    if opt.protect_5:
        f += ["--5trim_off"]
    if opt.discard_output:
        f += ["--discard"]
    if adapter if adapter is not None else opt.filter_adapter:
        f += ["--adapter"]
    if polyA:
        f += ["--polyA"]
    if artifact_file:
        f += ["--artifactFile", artifact_file]
    return f


def adapters_for(adapter: bool, polyA: bool, artifacts: Optional[Sequence[Tuple[str, str]]]) -> List[Tuple[str, str]]:
    """The adapter list the reference builds (options.cpp:576-694): built-ins, polyA, artifact file."""
    out: List[Tuple[str, str]] = []
    if adapter or artifacts:
        out += BUILTIN_ADAPTERS
    if polyA:
        out.append(POLYA_ADAPTER)
    if artifacts:
        out += list(artifacts)
    return out


def _read_matrix(path, ncol):
    if not os.path.exists(path):
        return np.zeros((0, ncol), dtype=np.uint64)
    rows = [list(map(int, ln.split("\t"))) for ln in open(path) if ln.strip()]
    return np.array(rows, dtype=np.uint64).reshape(-1, ncol)


def _read_qual_hist(path):
    reads = np.zeros(NUM_QUAL, dtype=np.uint64)
    bases = np.zeros(NUM_QUAL, dtype=np.uint64)
    with open(path) as fh:
        next(fh)
        for ln in fh:
            q, r, b = ln.split("\t")
            reads[int(q)] = int(r)
            bases[int(q)] = int(b)
    return reads, bases


def _read_base_content(path):
    comp = np.zeros((NUM_COMPOSITION, NUM_COMPOSITION_BIN), dtype=np.uint64)
    row = {"A": 0, "T": 1, "C": 2, "G": 3, "N": 4, "GC": 5}
    for ln in open(path):
        lab, pct, cnt = ln.split("\t")
        comp[row[lab], int(round(float(pct) * 100))] = int(cnt)
    return comp


def _read_length(path):
    vals = [0]
    for ln in open(path):
        l, c = ln.split("\t")
        assert int(l) == len(vals)
        vals.append(int(c))
    return np.array(vals, dtype=np.uint64)


def run_reference(r1: Optional[bytes] = None, r2: Optional[bytes] = None, unpaired: Optional[bytes] = None,
                  flags: Sequence[str] = (), threads: int = 1, artifacts: Optional[Sequence[Tuple[str, str]]] = None,
                  debug: bool = True, keep_dir: Optional[str] = None, timeout: int = 600) -> Dict:
    """Write inputs to a temp dir, run the reference with -t `threads`, return everything it wrote."""
    assert have_ref(), "oracle/_ref/FaQCs missing: run `make -C oracle ref`"
    tmp = keep_dir or tempfile.mkdtemp(prefix="faqcs_ref_")
    try:
        out = os.path.join(tmp, "out")
        cmd = [REF_BIN, "-d", out, "-t", str(threads)] + list(flags)
        if r1 is not None:
            open(os.path.join(tmp, "r1.fq"), "wb").write(bytes(r1))
            open(os.path.join(tmp, "r2.fq"), "wb").write(bytes(r2))
            cmd += ["-1", os.path.join(tmp, "r1.fq"), "-2", os.path.join(tmp, "r2.fq")]
        if unpaired is not None:
            open(os.path.join(tmp, "u.fq"), "wb").write(bytes(unpaired))
            cmd += ["-u", os.path.join(tmp, "u.fq")]
        if artifacts:
            with open(os.path.join(tmp, "artifacts.fa"), "w") as fh:
                for name, seq in artifacts:
                    fh.write(f">{name}\n{seq}\n")
            cmd += ["--artifactFile", os.path.join(tmp, "artifacts.fa")]
        if debug:
            cmd += ["--debug"]
        else:
            cmd += ["--trim_only"]
        # R is not installed: the reference pipes its plot script into a dead `sh -c R` and would die of
        # SIGPIPE now and then; the data files are complete by then, so let the write fail quietly instead
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout,
                           preexec_fn=lambda: signal.signal(signal.SIGPIPE, signal.SIG_IGN))
        res: Dict = {"returncode": p.returncode, "stderr": p.stderr.decode(errors="replace"), "cmd": cmd}

        def rd(name):
            path = os.path.join(out, name)
            return open(path, "rb").read() if os.path.exists(path) else None

        res["streams"] = [rd("QC.1.trimmed.fastq") or b"", rd("QC.2.trimmed.fastq") or b"",
                          rd("QC.unpaired.trimmed.fastq") or b"", rd("QC.discard.trimmed.fastq") or b""]
        st = rd("QC.stats.txt")
        res["stats_txt"] = st.decode() if st is not None else None
        if debug and p.returncode == 0:
            j = lambda n: os.path.join(out, n)
            res["pre_quality_matrix"] = _read_matrix(j("qa.QC.quality.matrix"), NUM_QUAL)
            res["post_quality_matrix"] = _read_matrix(j("QC.quality.matrix"), NUM_QUAL)
            res["pre_base_matrix"] = _read_matrix(j("qa.QC.base.matrix"), 5)
            res["post_base_matrix"] = _read_matrix(j("QC.base.matrix"), 5)
            res["pre_read_quality_hist"], res["pre_base_quality_hist"] = _read_qual_hist(j("qa.QC.for_qual_histogram.txt"))
            res["post_read_quality_hist"], res["post_base_quality_hist"] = _read_qual_hist(j("QC.for_qual_histogram.txt"))
            res["pre_composition"] = _read_base_content(j("qa.QC.base_content.txt"))
            res["post_composition"] = _read_base_content(j("QC.base_content.txt"))
            res["pre_length_hist"] = _read_length(j("qa.QC.length_count.txt"))
            res["post_length_hist"] = _read_length(j("QC.length_count.txt"))
            res["files"] = {n: open(j(n), "rb").read() for n in sorted(os.listdir(out))
                            if n.endswith((".matrix", ".txt"))}
        return res
    finally:
        if keep_dir is None:
            shutil.rmtree(tmp, ignore_errors=True)
