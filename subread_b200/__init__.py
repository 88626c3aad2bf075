"""subread_b200 -- B200-native seed-and-vote engine for Subread's read-mapping hot path.

The product is ``libsubread_cuda.so`` (C ABI in include/subread_cuda.h, CUDA sources in
subread_b200/csrc).  This package is the thin Python mirror used by tests and bench.py:
layouts, read batching and a ctypes binding.  There is no CPU fallback: every compute entry
point fails loudly when the CUDA library or a GPU is missing.
"""
from .layout import MAPPING_RESULT, SUBJUNC_RESULT, VoteParams, used_slots_equal  # noqa: F401
from .reads import ReadChunk, load_fastq, load_fastq_indexed, pack_reads, synthetic_reads  # noqa: F401
