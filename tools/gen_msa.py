#!/usr/bin/env python
"""Synthetic multiple alignments (groups) for the group-to-group DP tests and benchmarks.

An MSA is simulated column by column from a root sequence: every member copies, substitutes or
deletes each root column, and insertions open new columns that are gaps in every other member, so the
result is a true alignment with internal gaps (inex.dels) whose gap profile has several distinct
run lengths per column (Gfq::hetero > 1).  Written in the reference's native MSA format
(sample/pas/Multi_A: "<many> <len>\\t<name>" header, then ">name", residue lines and "/" per member)."""
import random

PROT = "ARNDCQEGHILKMFPSTWYV"
DNA = "ACGT"


def synth_msa(nmem, root_len, d0, d1, seed, dna=False, gapless=False):
    rng = random.Random(seed)
    alpha = DNA if dna else PROT
    root = [rng.choice(alpha) for _ in range(root_len)]
    cols = [[c] for c in root]          # each column: list of member characters so far
    rows = []
    # build member by member over a growing column list: columns = list of dicts
    columns = [{"root": c} for c in root]
    members = []
    for k in range(nmem):
        d = rng.uniform(d0, d1)
        out = {}                          # column object id -> char
        newcols = []
        i = 0
        seq_cols = []
        for col in columns:
            r = rng.random()
            base = col.get("root") or rng.choice(alpha)
            if gapless:
                ch = base if r > 0.7 * d else rng.choice(alpha)
            elif r < 0.70 * d:
                ch = rng.choice(alpha)
            elif r < 0.85 * d:
                ch = "-"
            else:
                ch = base
            col.setdefault("chars", {})[k] = ch
            newcols.append(col)
            if not gapless and rng.random() < 0.15 * d * 0.5:
                ins = {"root": None, "chars": {k: rng.choice(alpha)}}
                newcols.append(ins)
                # extend an insertion run sometimes
                while rng.random() < 0.4:
                    newcols.append({"root": None, "chars": {k: rng.choice(alpha)}})
        columns = newcols
    rows = []
    for k in range(nmem):
        rows.append("".join(col.get("chars", {}).get(k, "-") for col in columns))
    # drop all-gap columns
    keep = [j for j in range(len(columns)) if any(r[j] != "-" for r in rows)]
    rows = ["".join(r[j] for j in keep) for r in rows]
    # a member must not be empty
    rows = [r if any(c != "-" for c in r) else alpha[0] + r[1:] for r in rows]
    return rows


def split_family(rows, idx_a, idx_b):
    """Two sub-alignments of one family (as prrn's partitions are): members idx_a / idx_b with the
    columns that are all-gap inside the sub-alignment removed."""
    out = []
    for idx in (idx_a, idx_b):
        sub = [rows[i] for i in idx]
        keep = [j for j in range(len(sub[0])) if any(r[j] != "-" for r in sub)]
        out.append(["".join(r[j] for j in keep) for r in sub])
    return out


def write_native(path, rows, name="grp"):
    with open(path, "w") as f:
        f.write("%5d %5d\t%s\n" % (len(rows), len(rows[0]), name))
        for k, r in enumerate(rows):
            f.write(">%s_%d\n" % (name, k))
            for i in range(0, len(r), 60):
                f.write(r[i:i + 60] + "\n")
            f.write("/\n")


if __name__ == "__main__":
    import sys
    rows = synth_msa(int(sys.argv[1]), int(sys.argv[2]), 0.1, 0.5, int(sys.argv[3]))
    write_native(sys.argv[4], rows)
