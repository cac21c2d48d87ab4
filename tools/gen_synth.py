#!/usr/bin/env python
"""Fixed-seed synthetic sequence sets of the shapes BASELINE.json names.

A root sequence of the nominal length is drawn uniformly from the alphabet; each member is the root
mutated with a per-sequence divergence d ~ U(d0, d1): per site substitution w.p. 0.70 d, deletion
0.15 d, insertion (keep + one random letter) 0.15 d.  Deterministic for a given (seed, n, length).
"""
import argparse
import random

AA = "ARNDCQEGHILKMFPSTWYV"
NT = "ACGT"


def synth_set(n, length, d0, d1, seed, alphabet=AA, long_del=0.0):
    rng = random.Random(seed)
    root = [rng.choice(alphabet) for _ in range(length)]
    out = []
    for _ in range(n):
        d = rng.uniform(d0, d1)
        s = []
        skip = 0
        for c in root:
            if skip:
                skip -= 1
                continue
            r = rng.random()
            if r < 0.70 * d:
                s.append(rng.choice(alphabet))
            elif r < 0.85 * d:
                continue
            elif r < d:
                s.append(c)
                s.append(rng.choice(alphabet))
            else:
                s.append(c)
            if long_del and rng.random() < long_del / 20.0:
                skip = rng.randint(5, 40)
        if not s:
            s = [rng.choice(alphabet)]
        out.append("".join(s))
    return out


def write_fasta(path, seqs, prefix="s"):
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            f.write(">%s%d\n" % (prefix, i))
            for k in range(0, len(s), 60):
                f.write(s[k:k + 60] + "\n")


CONFIGS = {
    # name: (n, root length, d0, d1, seed, alphabet, long_del)
    "c2": (1000, 400, 0.1, 0.6, 1, AA, 0.0),
    "c3": (200, 500, 0.1, 0.6, 1, AA, 0.0),
    "c4": (100, 2000, 0.05, 0.35, 3, NT, 0.02),
    "c5a": (10000, 300, 0.1, 0.6, 5, AA, 0.0),
}


def config_set(name, n=None):
    nn, length, d0, d1, seed, alpha, ld = CONFIGS[name]
    seqs = synth_set(nn, length, d0, d1, seed, alpha, ld)
    return seqs[:n] if n else seqs


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", choices=sorted(CONFIGS))
    ap.add_argument("-n", type=int, default=0)
    ap.add_argument("--length", type=int, default=400)
    ap.add_argument("--d0", type=float, default=0.1)
    ap.add_argument("--d1", type=float, default=0.6)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--dna", action="store_true")
    ap.add_argument("out")
    a = ap.parse_args()
    if a.config:
        seqs = config_set(a.config, a.n or None)
    else:
        seqs = synth_set(a.n or 10, a.length, a.d0, a.d1, a.seed, NT if a.dna else AA)
    write_fasta(a.out, seqs)
