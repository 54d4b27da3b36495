"""Builds the corpus the fixed deflate code of the BGZF stage is fitted to and regenerates
mimick_b200/csrc/mmk_bgzf_table.h (like the other generators in this directory, committed next to what it makes).

    python tests/golden/make_bgzf_table.py

The corpus is FASTQ text of every barcode format written by the CPU oracle (test infrastructure) for two synthetic
haplotypes; the fitting itself and the bit-exact model of the coder live in scripts/bgzf_table.py.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import bgzf_table  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def sample_texts():
    rng = np.random.default_rng(2024)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    tmp = "/tmp/bgzf_table_inputs"
    os.makedirs(tmp, exist_ok=True)
    # a contig whose coordinates mix 6..8 digits, two haplotypes
    seqs = [acgt[rng.integers(0, 4, size=12_000_000)].tobytes() for _ in range(2)]
    paths = []
    for i, s in enumerate(seqs):
        p = os.path.join(tmp, f"h{i}.fa")
        with open(p, "wb") as fh:
            fh.write(b">chr7\n" + s + b"\n")
        paths.append(p)
    texts = []
    for fmt in ["haplotagging", "stlfr", "tellseq", "10x", "standard:haplotagging", "standard"]:
        res = orc.mimick_single(paths, fmt, seed=3, coverage=0.12, singletons=0.1)
        texts.append((fmt, res.r1, res.r2))
    return texts


if __name__ == "__main__":
    bgzf_table.fit(sample_texts())
