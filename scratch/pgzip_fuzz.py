"""Fuzz campaign for the parallel gzip reader (faqcs_b200/host/pgzip.hpp) on the CPU: gzip files of many shapes (data kinds,
compression levels and strategies, window sizes, flush points, members) through tests/host_io_harness (built as /tmp/hio);
what arrives must be the original bytes.   python scratch/pgzip_fuzz.py FIRST LAST"""
import os, subprocess, sys, zlib
import numpy as np
sys.path.insert(0, "/root/repo")
from faqcs_b200 import synth

HIO = "/tmp/hio"
base = bytes(synth.shotgun(30000, genome_len=20000).r1)          # repetitive FASTQ (deep coverage of a tiny genome)
rand_fq = bytes(synth.c2(30000).r1)
bad = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(seed)
    kind = seed % 6
    n = int(rng.integers(9_000_000, 16_000_000))
    if kind == 0:
        d = (base * 3)[:n]
    elif kind == 1:
        d = (rand_fq * 3)[:n]
    elif kind == 2:                       # long runs and far matches
        unit = rng.integers(65, 91, size=int(rng.integers(1, 40000)), dtype=np.uint8).tobytes()
        d = (unit * (n // len(unit) + 1))[:n]
    elif kind == 3:                       # small alphabet text with occasional binary bytes
        a = rng.choice(np.frombuffer(b"ACGTN\n@+I5", dtype=np.uint8), size=n)
        a[rng.integers(0, n, size=n // 5000)] = rng.integers(0, 256, size=n // 5000, dtype=np.uint8)
        d = a.tobytes()
    elif kind == 4:                       # mixture: compressible and incompressible stretches
        parts = []
        while sum(map(len, parts)) < n:
            parts.append(rng.integers(0, 256, size=int(rng.integers(1000, 300000)), dtype=np.uint8).tobytes() if rng.integers(0, 2) else base[:int(rng.integers(1000, 900000))])
        d = b"".join(parts)[:n]
    else:
        d = base[:n // 2] + rand_fq[:n // 2]
    level = int(rng.choice([1, 2, 4, 6, 9]))
    strategy = int(rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_RLE, zlib.Z_HUFFMAN_ONLY, zlib.Z_FIXED]))
    wbits = int(rng.choice([15, 15, 15, 12, 9]))
    members = int(rng.choice([1, 1, 2, 5]))
    flushes = int(rng.choice([0, 0, 3, 40]))
    z = bytearray()
    cuts = sorted(rng.integers(0, len(d), size=members - 1).tolist()) + [len(d)]
    at = 0
    for c in cuts:
        co = zlib.compressobj(level, zlib.DEFLATED, 16 + wbits, 8, strategy)
        seg = d[at:c]
        fl = sorted(rng.integers(0, max(1, len(seg)), size=flushes).tolist()) + [len(seg)]
        p = 0
        for f in fl:
            z += co.compress(seg[p:f])
            if f != len(seg):
                z += co.flush(int(rng.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH])))
            p = f
        z += co.flush()
        at = c
    small = len(z) < (4 << 20)            # the reader leaves small files to zlib unless told otherwise: shrink its units then
    env = dict(os.environ, FAQCS_B200_PGZIP_MIN="1000", FAQCS_B200_PGZIP_PIECE=str(max(64, len(z) // 16)), FAQCS_B200_PGZIP_SPAN=str(max(4096, len(z) // 3))) if small or seed % 2 else dict(os.environ)
    open("/tmp/pgzf.gz", "wb").write(bytes(z))
    cap = int(rng.choice([1500, 20000, 65536]))
    r = subprocess.run([HIO, "read", "/tmp/pgzf.gz", "/tmp/pgzf.out", str(cap)], capture_output=True, env=env)
    ok = r.returncode == 0 and open("/tmp/pgzf.out", "rb").read() == d and b"mode pgzip" in r.stdout
    if not ok:
        bad += 1
        print("seed", seed, "FAILED kind", kind, "level", level, "strategy", strategy, "wbits", wbits, "members", members, "flushes", flushes, "bytes", len(z), r.returncode,
              r.stderr.decode()[:200], flush=True)
print("seeds", sys.argv[1], "..", sys.argv[2], "failures", bad)
