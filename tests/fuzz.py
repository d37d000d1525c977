"""Random FASTQ records and option sets shared by the fuzz tests (oracle vs reference on CPU, CUDA vs oracle on GPU)."""
import numpy as np

from faqcs_b200.api import BUILTIN_ADAPTERS, POLYA_ADAPTER


def fuzz_reads(rng, n, in_off, paired_tag=None):
    """Random records exercising every per-base branch: lengths 1..420 (all phase-1 widths and the generic
    path), lower case, IUPAC letters, terminal / internal N runs, Q2 tails, '+name' third lines."""
    out = []
    for i in range(n):
        kind = int(rng.integers(0, 10))
        L = int(rng.integers(1, 40)) if kind == 0 else int(rng.integers(300, 421)) if kind == 1 else int(rng.integers(30, 200))
        al = "ACGT" if kind < 6 else "ACGTN" if kind < 8 else "ACGTNacgtnRYKM"
        s = "".join(rng.choice(list(al), size=L))
        if kind == 2:
            k = int(rng.integers(1, 6))
            s = ("N" * k + s + "N" * int(rng.integers(0, 4)))[:max(L, 1)]
        if kind == 3 and L > 20:
            p = int(rng.integers(0, L - 6))
            s = s[:p] + "N" * int(rng.integers(2, 6)) + s[p + 5:]
        if kind == 4:
            s = ("AC" * L)[:L] if i % 2 else "G" * L
        L = len(s)
        base = int(rng.integers(8, 41))
        q = np.clip(base - np.arange(L) * float(rng.uniform(0, 0.15)) + rng.normal(0, 3, size=L), 0, 41).astype(int)
        if kind in (5, 6) and L > 12:
            q[L - int(rng.integers(1, min(L, 70))):] = 2
        if kind == 7 and L > 12:
            q[:int(rng.integers(1, 10))] = 2
        qs = "".join(chr(in_off + int(x)) for x in q)
        tag = f" {paired_tag}:N:0:1" if paired_tag else ""
        out.append((f"@F{i}{tag}", s, qs))
    return out


def fuzz_bytes(recs, rng, eol):
    parts = []
    for h, s, q in recs:
        plus = "+" + h[1:] if rng.integers(0, 5) == 0 else "+"
        parts.append(f"{h}{eol}{s}{eol}{plus}{eol}{q}{eol}")
    return np.frombuffer("".join(parts).encode(), dtype=np.uint8)


def fuzz_options(rng, in_off, adapters=False):
    kw = dict(mode=int(rng.integers(0, 3)), quality=int(rng.choice([2, 5, 10, 20, 30])), trim_5=int(rng.choice([0, 0, 3, 40])),
              trim_3=int(rng.choice([0, 0, 5, 60])), min_read_length=int(rng.choice([1, 20, 50, 100])),
              max_num_poly_N=int(rng.choice([0, 1, 2, 4])), average_quality=float(rng.choice([0.0, 15.0, 27.5])),
              low_complexity_cutoff_ratio=float(rng.choice([0.85, 0.5, 0.3])), input_quality_offset=in_off,
              output_quality_offset=int(rng.choice([33, 33, 64])), replace_to_N_q=int(rng.choice([0, 0, 6, 20])),
              qc_only=bool(rng.integers(0, 6) == 0), protect_5=bool(rng.integers(0, 3) == 0), discard_output=bool(rng.integers(0, 2)))
    if adapters:
        kw.update(filter_adapter=True, adapters=list(BUILTIN_ADAPTERS) + [POLYA_ADAPTER], num_thread=int(rng.choice([1, 3])),
                  adapter_mismatch_rate=float(rng.choice([0.2, 0.1])))
    return kw


def adapter_fuzz_case(seed):
    """Adapter-centred case for the segment sweep: adapter sets of every shape (1..3 segments, more than 32 segments in all,
    IUPAC letters and gaps, one adapter beyond the sweep's 96 bases), reads that carry mutated pieces of them at either end
    or inside, short reads (threshold taken from the read), reads beyond 160 bases (several window passes), 'N' in the reads,
    mismatch rates from 0 to 1, the -t dependent threshold groups."""
    rng = np.random.default_rng(7000 + seed)
    bases = np.array(list("ACGT"))
    rnd = lambda n, al=bases: "".join(rng.choice(al, size=n))
    n_ad = int(rng.choice([1, 3, 12, 45]))
    adapters = []
    for j in range(n_ad):
        T = int(rng.choice([8, 20, 31, 32, 33, 40, 64, 65, 80, 96])) if j else int(rng.choice([33, 64, 97, 130]))
        s = rnd(T)
        if seed % 3 == 2 and j % 2:              # IUPAC / gap letters: the general (non bit-select) form of the sweep
            s = list(s)
            for p in rng.integers(0, T, size=max(1, T // 10)):
                s[p] = str(rng.choice(list("RYKMSWBDHVN-")))
            s = "".join(s)
        adapters.append((f"A{j}", s))
    if seed % 4 == 1:
        adapters += list(BUILTIN_ADAPTERS) + [POLYA_ADAPTER]
    n = 600
    recs = []
    for i in range(n):
        kind = int(rng.integers(0, 8))
        L = int(rng.integers(1, 40)) if kind == 0 else int(rng.integers(161, 330)) if kind == 1 else int(rng.integers(40, 161))
        s = list(rnd(L, np.array(list("ACGTN")) if kind == 2 else bases))
        if kind >= 3 and adapters:
            a = list(adapters[int(rng.integers(0, len(adapters)))][1].replace("-", "A"))
            a = [c if c in "ACGT" else "A" for c in a]
            for p in rng.integers(0, len(a), size=int(rng.integers(0, max(1, len(a) // 5)))):
                a[p] = str(rng.choice(bases))
            cut = int(rng.integers(0, len(a)))
            piece = a[cut:] if kind == 3 else a[:len(a) - cut] if kind == 4 else a
            pos = 0 if kind == 3 else max(0, L - len(piece)) if kind == 4 else int(rng.integers(0, max(1, L - len(piece) + 1)))
            for k, c in enumerate(piece):
                if pos + k < L:
                    s[pos + k] = c
        q = "".join(chr(33 + int(x)) for x in rng.integers(2, 41, size=L))
        recs.append((f"@A{i}", "".join(s), q))
    kw = dict(filter_adapter=True, adapters=adapters, num_thread=int(rng.choice([0, 1, 3, 7])),
              adapter_mismatch_rate=float(rng.choice([0.2, 0.2, 0.1, 0.0, 0.5, 1.0])), min_read_length=int(rng.choice([1, 30])),
              quality=int(rng.choice([2, 5, 20])), input_quality_offset=33, discard_output=True)
    return recs, kw
