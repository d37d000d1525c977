"""Command-line fuzz campaign: random option sets x random reads through the reference binary and through faqcs_b200; every
file both leave must be byte-identical, and a run the reference refuses must be refused.  python scratch/cli_fuzz.py FIRST LAST"""
import gzip, os, shutil, signal, subprocess, sys, tempfile
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import refcli
from faqcs_b200.api import Options
from fuzz import fuzz_bytes, fuzz_options, fuzz_reads
from test_host_io import bgzf_bytes
CLI = "/root/repo/faqcs_b200/host/faqcs_b200"
bad = 0
refused = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(9000 + seed)
    in_off = 64 if seed % 4 == 3 else 33
    paired = seed % 2 == 0
    eol = "\r\n" if seed % 5 == 4 else "\n"
    n = int(rng.choice([300, 2000]))
    r1 = fuzz_bytes(fuzz_reads(rng, n, in_off, "1" if paired else None), rng, eol)
    r2 = fuzz_bytes(fuzz_reads(rng, n, in_off, "2"), rng, eol) if paired else None
    kw = fuzz_options(rng, in_off, adapters=False)
    adapter = seed % 3 == 1
    polyA = seed % 6 == 1
    kw.pop("input_quality_offset")
    if seed % 7 == 0:
        kw["input_quality_offset"] = in_off               # explicit --ascii, else autodetection
    opt = Options(**kw)
    threads = int(rng.choice([1, 2, 5]))
    flags = refcli.flags_for(opt, polyA=polyA, adapter=adapter)
    if seed % 8 == 5:
        flags += ["--kmer_rarefaction", "--split_size", str(int(rng.choice([200, 1000]))), "--subset", "2"]
    tmp = tempfile.mkdtemp(prefix="faqcs_clifuzz_")
    # host pipeline variety (faqcs_b200 only; the reference reads the same files): input as plain / gzip / blocked gzip,
    # small batches, two contexts on the device, blocked gzip output
    def put(path, data, how):
        d = bytes(data)
        open(path, "wb").write(d if how == 0 else gzip.compress(d, 1) if how == 1 else bgzf_bytes(d, block=int(rng.choice([900, 0xff00]))))
    extra = []
    if seed % 2:
        extra += ["--batch_mb", "1"]
    if seed % 5 == 2:
        extra += ["--devices", "0,0"]
    gz_out = seed % 4 == 3 and not kw.get("qc_only")
    if gz_out:
        extra += ["--gz_out"]
    try:
        args = []
        if paired:
            put(f"{tmp}/r1.fq", r1, int(rng.integers(0, 3))); put(f"{tmp}/r2.fq", r2, int(rng.integers(0, 3)))
            args = ["-1", f"{tmp}/r1.fq", "-2", f"{tmp}/r2.fq"]
        else:
            put(f"{tmp}/u.fq", r1, int(rng.integers(0, 3)))
            args = ["-u", f"{tmp}/u.fq"]
        outs, rcs = {}, {}
        for tag, exe in (("ref", refcli.REF_BIN), ("gpu", CLI)):
            out = f"{tmp}/{tag}"
            # odd seeds: the parallel gzip reader on these small files too (its units shrunk)
            env = dict(os.environ, FAQCS_B200_PGZIP_MIN="1000", FAQCS_B200_PGZIP_PIECE="20000", FAQCS_B200_PGZIP_SPAN="150000") if tag == "gpu" and seed % 2 else None
            p = subprocess.run([exe, "-d", out, "-t", str(threads), "--debug"] + args + flags + (extra if tag == "gpu" else []), stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env,
                               preexec_fn=lambda: signal.signal(signal.SIGPIPE, signal.SIG_IGN))
            rcs[tag] = p.returncode
            outs[tag] = {f: open(f"{out}/{f}", "rb").read() for f in sorted(os.listdir(out)) if not f.endswith(".pdf")} if os.path.isdir(out) else {}
            if tag == "gpu" and gz_out:
                outs[tag] = {(f[:-3] if f.endswith(".fastq.gz") else f): (gzip.decompress(v) if f.endswith(".fastq.gz") else v) for f, v in outs[tag].items()}
            if tag == "gpu" and p.returncode != 0:
                gpu_err = p.stderr.decode(errors="replace")[-300:]
        if (rcs["ref"] == 0) != (rcs["gpu"] == 0):
            bad += 1
            print("seed", seed, "exit codes differ", rcs, flags, flush=True)
            continue
        if rcs["ref"] != 0:
            refused += 1
            continue
        names = sorted(set(outs["ref"]) | set(outs["gpu"]))
        diff = [f for f in names if outs["ref"].get(f) != outs["gpu"].get(f)]
        if diff:
            bad += 1
            print("seed", seed, "files differ:", diff, " ".join(flags), "paired" if paired else "single", "t", threads, flush=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
print("seeds", sys.argv[1], "..", sys.argv[2], "failures", bad, "(runs the reference refused:", refused, ")")
