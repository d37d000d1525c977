"""Option handling of the drop-in command line against the reference binary, without a GPU: command lines that both programs
must refuse before any read is processed (options.cpp:181-574) -- same exit code, same first message."""
import os
import re
import subprocess

import pytest

import refcli

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "faqcs_b200", "host", "faqcs_b200")
pytestmark = [pytest.mark.skipif(not refcli.have_ref(), reason="reference binary not built"),
              pytest.mark.skipif(not os.path.exists(CLI), reason="CLI not built")]

CASES = {
    "no_arguments": [],
    "only_read_one": ["-1", "X", "-d", "OUT"],
    "only_read_two": ["-2", "X", "-d", "OUT"],
    "no_input": ["-d", "OUT", "--qc_only"],
    "lc_out_of_range": ["-u", "X", "-d", "OUT", "--lc", "1.5"],
    "rate_out_of_range": ["-u", "X", "-d", "OUT", "--rate", "2"],
    "split_size_zero": ["-u", "X", "-d", "OUT", "--split_size", "0"],
    "subset_zero": ["-u", "X", "-d", "OUT", "--subset", "0"],
    "kmer_flag_is_ambiguous": ["-u", "X", "-d", "OUT", "-m", "20"],          # -m abbreviates --mode / --min_L under getopt_long_only
    "unknown_flag": ["-u", "X", "-d", "OUT", "--bogus"],
    "version": ["--version"],
    "help": ["-h"],
    "q_out_of_bounds": ["-u", "X", "-d", "OUT", "-q", "200"],
    "ascii_out_of_bounds": ["-u", "X", "-d", "OUT", "--ascii", "300"],
    "missing_input_file": ["-u", "X", "-d", "OUT"],                            # X does not exist: refused when the pass opens it
}


def run(exe, args, tmp):
    args = [a.replace("OUT", os.path.join(tmp, "out")).replace("X", os.path.join(tmp, "missing.fq")) if a in ("OUT", "X") else a for a in args]
    p = subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    first = (p.stderr.decode(errors="replace").splitlines() or [""])[0]
    first = re.sub(r"/\S*/(\w+\.cpp)", r"\1", first)                          # __FILE__ of the reference build carries its path
    first = first.replace(" (faqcs_b200: GPU trim/filter/statistics path)", "").replace(tmp, "TMP")
    return p.returncode, first


@pytest.mark.parametrize("name", list(CASES), ids=list(CASES))
def test_refused_command_lines(name, tmp_path):
    if name == "missing_input_file":
        # this one reaches fq_create in faqcs_b200 (a device is needed to get that far): compared on the GPU box only
        try:
            import torch
            if not torch.cuda.is_available():
                pytest.skip("needs a device: the input file is opened after the context is created")
        except Exception:
            pytest.skip("torch missing")
    a = run(refcli.REF_BIN, CASES[name], str(tmp_path))
    b = run(CLI, CASES[name], str(tmp_path))
    assert a == b
