"""Residue codes of the reference (src/cmn.h:110-112, Appendix A of SURVEY.md).

protein: NIL=0, UNP(gap)=1, AMB=2, ALA=3 ... VAL=22, ASX=23, GLX=24, TRM=25
nucleotide: nil=0, gap=1, then 1 + 4-bit set (A=1, C=2, G=4, T=8): A=2, C=3, G=5, T=9, N=16
"""
import numpy as np

_AA_ORDER = "ARNDCQEGHILKMFPSTWYV"
PROT_DIM = 25  # ASIMD (seq.h:81)
NUC_DIM = 17   # NSIMD (seq.h:80)

_prot = np.full(256, 2, dtype=np.uint8)  # unknown letters -> AMB
for _i, _c in enumerate(_AA_ORDER):
    _prot[ord(_c)] = 3 + _i
    _prot[ord(_c.lower())] = 3 + _i
for _c, _v in (("B", 23), ("Z", 24), ("X", 2), ("-", 1), ("*", 25), ("U", 24), ("J", 2), ("O", 2)):
    _prot[ord(_c)] = _v
    _prot[ord(_c.lower())] = _v

_NUC_BITS = {"A": 1, "C": 2, "G": 4, "T": 8, "U": 8, "M": 3, "R": 5, "S": 6, "V": 7, "W": 9,
             "Y": 10, "H": 11, "K": 12, "D": 13, "B": 14, "N": 15}
_nuc = np.full(256, 16, dtype=np.uint8)
for _c, _b in _NUC_BITS.items():
    _nuc[ord(_c)] = 1 + _b
    _nuc[ord(_c.lower())] = 1 + _b
_nuc[ord("-")] = 1


def encode_protein(s):
    return _prot[np.frombuffer(s.encode("ascii"), dtype=np.uint8)]


def encode_dna(s):
    return _nuc[np.frombuffer(s.encode("ascii"), dtype=np.uint8)]


def read_fasta(path):
    names, seqs, cur = [], [], []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            if line[0] == ">":
                if names:
                    seqs.append("".join(cur))
                names.append(line[1:].split()[0])
                cur = []
            else:
                cur.append(line)
    if names:
        seqs.append("".join(cur))
    return names, seqs


def pack(encoded):
    """Concatenate encoded sequences -> (residues u8, offsets i64[n+1], lens i32[n])."""
    lens = np.array([len(e) for e in encoded], dtype=np.int32)
    offs = np.zeros(len(encoded) + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    res = np.concatenate(encoded).astype(np.uint8) if len(encoded) else np.zeros(0, np.uint8)
    return res, offs, lens
