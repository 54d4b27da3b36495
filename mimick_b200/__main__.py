"""`python -m mimick_b200` — the reference's command line (mimick/__main__.py:36-90) over the B200 library.

Same options, short flags, defaults and validation as the reference's `mimick` console script; instead of
spawning `julia -e 'using MimickLinkedReads; mimick(...)'` the parsed options go straight to
`mimick_b200.mimick(...)` (C ABI → CUDA).  Additions: `--device`, `--plain` (uncompressed .fq instead of BGZF
.fq.gz) and `--rank/--world` for one process per GPU.  `--threads` is accepted and ignored (the reference uses it
for Julia threads over samples; here samples are spread over ranks).
"""
from __future__ import annotations

import argparse
import os
import sys

FORMATS = ["10x", "stlfr", "standard", "standard:haplotagging", "standard:stlfr", "haplotagging", "tellseq"]


def _read_lengths(value: str):
    """cli_classes.py:ReadLengths — two integers >= 10 separated by a comma"""
    try:
        r1, r2 = value.split(",")
    except ValueError:
        raise argparse.ArgumentTypeError(f"{value} is not in int,int format")
    out = []
    for name, v in (("R1", r1), ("R2", r2)):
        try:
            n = int(v)
        except ValueError:
            raise argparse.ArgumentTypeError(f"{v} is not an integer.")
        if n < 10:
            raise argparse.ArgumentTypeError(f"{name} reads must be at least 10 bp. (input: {n})")
        out.append(n)
    return out


def _ranged(kind, lo=None, hi=None, clamp=False):
    def conv(value):
        try:
            v = kind(value)
        except ValueError:
            raise argparse.ArgumentTypeError(f"{value!r} is not a valid {kind.__name__}")
        if clamp:
            v = lo if lo is not None and v < lo else v
            v = hi if hi is not None and v > hi else v
        elif (lo is not None and v < lo) or (hi is not None and v > hi):
            raise argparse.ArgumentTypeError(f"{v} is not in the range {lo}..{hi if hi is not None else ''}")
        return v
    return conv


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="mimick", description="Simulate linked-read FASTQ data for one or many individuals",
                                 epilog="Documentation: https://pdimens.github.io/mimick/")
    ap.add_argument("--version", action="version", version="mimick (mimick_b200)")
    ap.add_argument("-c", "--circular", action="store_true", help="contigs are circular/prokaryotic")
    ap.add_argument("-o", "--output-prefix", default="simulated/", help="output file prefix")
    ap.add_argument("-q", "--quiet", action="store_true", help="toggle to hide progress bar")
    ap.add_argument("-m", "--mask-n", action="store_true", help="Replace Ns in the fasta file(s) with random ATCG bases")
    ap.add_argument("-t", "--threads", type=_ranged(int, 2), default=2, help="accepted for compatibility (ignored)")
    ap.add_argument("-s", "--seed", type=_ranged(int, -1, None, clamp=True), default=-1, help="random seed for simulation")
    ap.add_argument("-f", "--format", dest="fmt", default="standard:haplotagging", type=str.lower, choices=FORMATS,
                    help="FASTQ output format")
    ap.add_argument("-g", "--genomic-coverage", type=_ranged(float, 0.05), default=30.0,
                    help="mean coverage (depth) target for simulated data")
    ap.add_argument("-i", "--insert-size", type=_ranged(int, 0), default=500, help="outer distance between the two read ends in bp")
    ap.add_argument("-d", "--insert-stdev", type=_ranged(int, 0), default=50, help="standard deviation for --insert-size")
    ap.add_argument("-A", "--molecule-attempts", type=_ranged(int, 5), default=300,
                    help="how many tries to create a molecule with <70%% ambiguous bases")
    ap.add_argument("-C", "--molecule-coverage", type=_ranged(float, 0.00001), default=0.2,
                    help="mean percent coverage per molecule if <1, else mean number of reads per molecule")
    ap.add_argument("-L", "--molecule-length", type=_ranged(int, 300), default=80000, help="mean length of molecules in bp")
    ap.add_argument("-N", "--molecules-per", type=int, default=2,
                    help="mean number of unrelated molecules per barcode per chromosome (negative: fixed number)")
    ap.add_argument("-l", "--read-lengths", type=_read_lengths, default=[150, 150], help="length of R1,R2 sequences in bp")
    ap.add_argument("-S", "--singletons", type=_ranged(float, 0, 1), default=0, help="proportion of barcodes with one read pair")
    ap.add_argument("-v", "--vcf", help="VCF-formatted file containing genotypes from which to create per-sample haplotypes")
    ap.add_argument("--device", type=int, default=0, help="CUDA device")
    ap.add_argument("--plain", action="store_true", help="write uncompressed .fq instead of BGZF .fq.gz")
    ap.add_argument("--rank", type=int, default=0, help="this process's part of a multi-GPU run")
    ap.add_argument("--world", type=int, default=1, help="number of processes (one per GPU)")
    ap.add_argument("fasta", nargs="+", help="haplotype FASTA file(s)")
    return ap


def plan_call(argv):
    """argv → (positional args, keyword args) of mimick_b200.mimick — the mapping of __main__.py:73-83."""
    ns = build_parser().parse_args(argv)
    if sum(ns.read_lengths) > ns.molecule_length:
        raise SystemExit(f"ERROR: The sum of the read lengths {ns.read_lengths} cannot exceed the average molecule length "
                         f"({ns.molecule_length}).")
    fmt = "tenx" if ns.fmt == "10x" else ns.fmt
    kw = dict(coverage=ns.genomic_coverage, n_molecules=ns.molecules_per, mol_coverage=ns.molecule_coverage,
              mol_length=ns.molecule_length, insert_length=ns.insert_size, insert_stdev=ns.insert_stdev,
              read_length=ns.read_lengths, singletons=ns.singletons, circular=ns.circular, attempts=ns.molecule_attempts,
              seed=ns.seed, mask=ns.mask_n, quiet=ns.quiet, device=ns.device, compress=not ns.plain, rank=ns.rank, world=ns.world)
    # click resolves -o with resolve_path=True (mimick/__main__.py:37): the default "simulated/" becomes <cwd>/simulated, so the
    # files are simulated.R1.fq.gz (not hidden files inside simulated/), and the VCF mode's outdir is the resolved directory
    out = os.path.realpath(ns.output_prefix)
    if ns.vcf:
        return (ns.fasta[0], ns.vcf, fmt), dict(kw, outdir=out)
    return (list(ns.fasta), fmt), dict(kw, prefix=out)


def main(argv=None) -> int:
    args, kw = plan_call(sys.argv[1:] if argv is None else argv)
    from . import MimickError, mimick
    try:
        outs = mimick(*args, **kw)
    except MimickError as e:  # the reference exits 1 on any Julia error (__main__.py:90-102)
        print(f"ERROR: {e}", file=sys.stderr)
        return 1
    if not kw["quiet"]:
        for o in outs:
            print(f"{o.name or 'sample'}: {o.n_pairs} read pairs, {o.n_molecules} molecules -> {o.r1_path}, {o.r2_path} (seed {o.seed})")
    return 0


if __name__ == "__main__":
    sys.exit(main())
