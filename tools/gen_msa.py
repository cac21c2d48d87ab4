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


def synth_msa(nmem, root_len, d0, d1, seed, dna=False, gapless=False, indel_events=3.0):
    """Tree-structured family: member k descends from a random earlier member (member 0 from the root).
    A child copies its parent's row, substitutes residues with its own divergence d ~ U(d0, d1), deletes a
    few contiguous runs and inserts a few runs (new columns, gaps in every non-descendant), so gap runs
    are shared along lineages and the column count grows slowly (a few columns per member)."""
    rng = random.Random(seed)
    alpha = DNA if dna else PROT
    ncol = root_len
    rows = []                       # each row: dict column-id -> char; columns kept in an ordered list
    order = list(range(root_len))   # column ids in alignment order
    next_id = root_len
    root = {c: rng.choice(alpha) for c in order}
    for k in range(nmem):
        parent = root if k == 0 else rows[rng.randrange(k)]
        d = rng.uniform(d0, d1)
        child = {}
        for c in order:
            ch = parent.get(c)
            if ch is None:
                continue
            child[c] = rng.choice(alpha) if rng.random() < 0.7 * d else ch
        if not gapless:
            present = [c for c in order if c in child]
            nev = 0
            while rng.random() < indel_events * d / (1 + indel_events * d) and nev < 8:
                nev += 1
                if rng.random() < 0.5 and len(present) > 20:        # deletion of a contiguous run
                    ln = 1
                    while rng.random() < 0.6 and ln < 12:
                        ln += 1
                    st = rng.randrange(1, max(2, len(present) - ln - 1))
                    for c in present[st:st + ln]:
                        child.pop(c, None)
                    present = [c for c in order if c in child]
                else:                                               # insertion of a run of new columns
                    ln = 1
                    while rng.random() < 0.5 and ln < 8:
                        ln += 1
                    anchor = present[rng.randrange(len(present))]
                    pos = order.index(anchor) + 1
                    ids = list(range(next_id, next_id + ln))
                    next_id += ln
                    order[pos:pos] = ids
                    for c in ids:
                        child[c] = rng.choice(alpha)
                    present = [c for c in order if c in child]
        rows.append(child)
    out = ["".join(r.get(c, "-") for c in order) for r in rows]
    keep = [j for j in range(len(order)) if any(r[j] != "-" for r in out)]
    return ["".join(r[j] for j in keep) for r in out]


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
