"""Golden vectors produced by the unmodified reference binary (tests/golden/make_golden.py), replayed
against the CPU oracle (CPU suite) and the CUDA path (-m gpu).  They keep parity pinned where
/root/reference and oracle/_ref/FaQCs are absent."""
import ast
import glob
import os

import numpy as np
import pytest

from faqcs_b200.api import Engine, Options
from oracle_binding import OracleEngine
from parity import MATRIX_FIELDS, assert_matches_reference, run_engine

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def load(path):
    z = np.load(path)
    okw = ast.literal_eval(bytes(z["options"]).decode())
    adapters = ast.literal_eval(bytes(z["adapters"]).decode())
    opt = Options(**okw)
    opt.adapters = [tuple(a) for a in adapters]
    if adapters:
        opt.filter_adapter = True
    ref = {"returncode": 0, "stderr": "", "stats_txt": bytes(z["stats_txt"]).decode()}
    if "stream0" in z.files:
        ref["streams"] = [bytes(z[f"stream{i}"]) for i in range(4)]
    else:       # large cases: the reference's streams are committed as (length, sha256)
        ref["stream_hashes"] = [(int(z[f"stream{i}_len"][0]), bytes(z[f"stream{i}_sha256"])) for i in range(4)]
    for f in MATRIX_FIELDS:
        ref[f] = z[f]
    gen = bytes(z["generator"]).decode() if "generator" in z.files else ""
    if gen:     # input regenerated from the seeded generator (too large to commit)
        from faqcs_b200 import synth
        w = eval(gen, {"synth": synth})
        return w.r1, w.r2, opt, ref
    r2 = z["r2"] if bool(z["paired"][0]) else None
    return z["r1"], r2, opt, ref


# BASELINE-size case (2 M pairs): GPU only -- the oracle needs minutes for it and is pinned by the other fixtures
CPU_FIXTURES = [p for p in FIXTURES if "2m_pairs" not in p]


def test_fixtures_exist():
    assert len(FIXTURES) >= 9


@pytest.mark.parametrize("path", CPU_FIXTURES, ids=[os.path.basename(p)[:-4] for p in CPU_FIXTURES])
def test_oracle_reproduces_reference_golden(path):
    r1, r2, opt, ref = load(path)
    with OracleEngine(opt) as eng:
        streams, _ = run_engine(eng, r1, r2)
        assert_matches_reference(ref, streams, eng.stats(), opt, opt.adapters)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_cuda_reproduces_reference_golden(path):
    r1, r2, opt, ref = load(path)
    with Engine(opt) as eng:
        streams, _ = run_engine(eng, r1, r2)
        assert_matches_reference(ref, streams, eng.stats(), opt, opt.adapters)
