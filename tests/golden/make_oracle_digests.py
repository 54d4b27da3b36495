"""Freezes the CPU oracle: SHA-256 of its R1 / R2 text, molecule table and insert breakpoints for a set of seeded cases,
written to tests/golden/oracle_digests.json.

    python tests/golden/make_oracle_digests.py          # rewrites the file (only after a DELIBERATE change of the oracle
                                                         # or of include/mmk_rng.h; say so in the commit)

The GPU tests compare the CUDA path with the oracle run live, and both sides share include/mmk_rng.h (Philox and the
variate transforms), so a change there moves both at once and no GPU test notices.  tests/test_oracle_kat.py::
test_oracle_output_is_frozen compares these digests on the CPU: any drift of the oracle's bytes — through its own code or
through the shared header — fails there first.  The digests pin this repo's oracle to itself over time; they are NOT
reference outputs (the reference ships none and Julia cannot run here: parity stays unpinned upstream).
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle as O  # noqa: E402
from helpers import random_bases, write_fasta  # noqa: E402

OUT = os.path.join(HERE, "oracle_digests.json")
FIX = os.path.join(HERE, "ref_fixtures")

# (name, format, keyword arguments of the simulation)
SINGLE = [
    ("haplotagging_defaults", "haplotagging", dict(seed=1, coverage=6, attempts=300, singletons=0.2)),
    ("stlfr_circular_literal", "stlfr", dict(seed=2, coverage=8, attempts=300, singletons=0.0, circular=True, mol_length=15000)),
    ("stlfr_circular_strict", "stlfr", dict(seed=2, coverage=8, attempts=300, singletons=0.0, circular=True, mol_length=15000, strict_wrap=True)),
    ("tenx_fixed_nmol", "10x", dict(seed=3, coverage=5, attempts=300, singletons=0.5, n_molecules=-2, mol_length=30000)),
    ("tellseq_masked_n", "tellseq", dict(seed=4, coverage=5, attempts=300, singletons=0.1, mask=True)),
    ("standard_stlfr_short_reads", "standard:stlfr", dict(seed=5, coverage=4, attempts=300, singletons=0.7, read_length=(75, 101))),
    # truncated(Normal(150, 50), lower = 200) keeps 16 % of the mass: Robert's exponential proposal (Distributions.jl randnt)
    ("haplotagging_robert_inserts", "haplotagging", dict(seed=6, coverage=4, attempts=300, singletons=0.0, insert_length=150, insert_stdev=50,
                                                        mol_coverage=0.05)),
    ("haplotagging_error_model", "haplotagging", dict(seed=7, coverage=4, attempts=300, singletons=0.3, error_rate=(0.002, 0.02),
                                                     indel_error=(0.001, 0.0015))),
]


def inputs(tmp):
    rng = np.random.default_rng(20261020)
    p1, p2 = os.path.join(tmp, "a.fa"), os.path.join(tmp, "b.fa")
    write_fasta(p1, [("chrA", random_bases(rng, 70_000)), ("chrB", random_bases(rng, 12_000, b"ACGTN")), ("tiny", random_bases(rng, 1_004))])
    write_fasta(p2, [("chrA", random_bases(rng, 70_200)), ("chrB", random_bases(rng, 11_500)), ("tiny", random_bases(rng, 1_004))])
    p3 = os.path.join(tmp, "c.fa")     # without the 1 004-base contig (200-base inserts cannot fit 5 to a molecule there: the reference errors)
    write_fasta(p3, [("chrA", random_bases(rng, 70_000)), ("chrB", random_bases(rng, 12_000))])
    return [p1, p2], [p3]


def digest(res):
    h = lambda b: hashlib.sha256(b).hexdigest()
    m = res.molecules
    fields = ("group", "slot", "schema", "start", "molsize", "n_frags", "first_read", "barcode_index")
    return dict(r1=h(res.r1), r2=h(res.r2), r1_bytes=len(res.r1), pairs=int(len(res.insert_start)), molecules=int(len(m)),
                plan=h(b"".join(np.ascontiguousarray(m[f]).astype("<i8").tobytes() for f in fields)),
                inserts=h(np.ascontiguousarray(res.insert_start).astype("<i8").tobytes() + np.ascontiguousarray(res.insert_stop).astype("<i8").tobytes()))


def compute():
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        fastas, one = inputs(tmp)
        for name, fmt, kw in SINGLE:
            try:
                out[name] = digest(O.mimick_single(one if "robert" in name else fastas, fmt, **kw))
            except O.OracleError as e:
                raise SystemExit(f"{name}: {e}")
    # the reference's own fixtures, FASTA + VCF mode, two of the seven samples (src/Main.jl:88-131)
    multi = O.mimick_multi(os.path.join(FIX, "fasta.fa.gz"), os.path.join(FIX, "test.vcf.gz"), "haplotagging", seed=8, coverage=1,
                           attempts=300, singletons=0.2, samples=[0, 5])
    for sname, res in sorted(multi.items()):
        out["vcf_fixture_" + sname] = digest(res)
    return out


if __name__ == "__main__":
    d = compute()
    with open(OUT, "w") as fh:
        json.dump(d, fh, indent=1, sort_keys=True)
        fh.write("\n")
    print(OUT, len(d), "cases,", sum(v["pairs"] for v in d.values()), "read pairs")
