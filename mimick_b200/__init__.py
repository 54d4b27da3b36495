"""mimick_b200 — B200-native simulation core for the linked-read simulator pdimens/mimick.

The product is the C ABI in include/mimick_b200.h (libmimick_b200.so: hand-written sm_100a CUDA
kernels + a C++ host driver).  This package is the thin Python host above it that mirrors the
reference's `mimick(...)` API; importing it requires the built library — there is no CPU path.
"""
from . import _lib
from ._lib import MimickError
from .api import SampleOutput, Simulator, Vcf, bgzf_eof, interperet_format, mimick
from .dist import RankCounts, ShardLayout, exchange_counts, layout_from_counts, samples_for_rank

__all__ = ["mimick", "Simulator", "Vcf", "MimickError", "SampleOutput", "interperet_format", "bgzf_eof"]

_lib.load()  # fail loudly at import time if libmimick_b200.so is missing
