"""Plain-old-data layouts shared by the C-ABI library, the oracle binding and the tests.

These mirror the reference's structs byte for byte so that buffers can be handed to the
reference's own host code unchanged:

* ``MAPPING_RESULT`` -- ``mapping_result_t``  (reference src/core.h:350-370), 68 bytes
* ``SUBJUNC_RESULT`` -- ``subjunc_result_t``  (reference src/core.h:397-410), 16 bytes
* ``VoteParams``     -- the scalars ``do_voting`` / ``process_voting_junction_PE_topK`` read from
  ``configuration_t`` (reference src/core.h:128-253); same field order as ``srcuda_params``
  in include/subread_cuda.h.
"""
import ctypes

import numpy as np

MAPPING_RESULT = np.dtype(
    [
        ("selected_position", "<u4"),
        ("result_flags", "<i2"),
        ("read_length", "<i2"),
        ("selected_votes", "<i2"),
        ("used_subreads_in_vote", "<i2"),
        ("noninformative_subreads_in_vote", "u1"),
        ("indels_in_confident_coverage", "i1"),
        ("is_fully_covered", "i1"),
        ("_pad0", "u1"),
        ("selected_indel_record", "<i2", (22,)),
        ("confident_coverage_start", "<u2"),
        ("confident_coverage_end", "<u2"),
        ("subread_quality", "<i2"),
        ("_pad1", "<u2"),
    ]
)
assert MAPPING_RESULT.itemsize == 68

SUBJUNC_RESULT = np.dtype(
    [
        ("split_point", "<i2"),
        ("minor_votes", "<i2"),
        ("double_indel_offset", "i1"),
        ("indel_at_junction", "i1"),
        ("small_side_increasing_coordinate", "i1"),
        ("large_side_increasing_coordinate", "i1"),
        ("minor_position", "<u4"),
        ("minor_coverage_start", "<u2"),
        ("minor_coverage_end", "<u2"),
    ]
)
assert SUBJUNC_RESULT.itemsize == 16

BIG_MARGIN_U16_PER_READ = 27  # bigtable_cached_result_t.big_margin_data, reference src/core.h:453
CORE_IS_NEGATIVE_STRAND = 8  # reference src/core.h:258
MAX_READ_LENGTH = 1210  # reference src/subread.h:73


class VoteParams(ctypes.Structure):
    _fields_ = [
        ("total_subreads", ctypes.c_int32),
        ("top_scores", ctypes.c_int32),
        ("max_vote_simples", ctypes.c_int32),
        ("max_vote_combinations", ctypes.c_int32),
        ("max_vote_number_cutoff", ctypes.c_int32),
        ("multi_best_reads", ctypes.c_int32),
        ("minimum_subread_for_first_read", ctypes.c_int32),
        ("minimum_subread_for_second_read", ctypes.c_int32),
        ("max_indel_length", ctypes.c_int32),
        ("minimum_pair_distance", ctypes.c_int32),
        ("maximum_pair_distance", ctypes.c_int32),
        ("is_first_read_reversed", ctypes.c_int32),
        ("is_second_read_reversed", ctypes.c_int32),
        ("do_breakpoint_detection", ctypes.c_int32),
        ("do_big_margin_filtering_for_junctions", ctypes.c_int32),
        ("big_margin_record_size", ctypes.c_int32),
        ("maximum_intron_length", ctypes.c_int32),
        ("limited_tree_scan", ctypes.c_int32),
        ("check_donor_at_junctions", ctypes.c_int32),
        ("more_accurate_fusions", ctypes.c_int32),
        ("max_insertion_at_junctions", ctypes.c_int32),
        ("prefer_donor_receptor_junctions", ctypes.c_int32),
        ("reserved", ctypes.c_int32 * 2),
    ]

    @classmethod
    def defaults(cls, is_paired: bool, is_subjunc: bool = False) -> "VoteParams":
        """subread-align / subjunc defaults after ``load_global_context``'s overrides
        (reference src/core-indel.c:4399-4538, src/core.c:4075-4084,
        src/core-interface-subjunc.c:264-284)."""
        p = cls()
        p.total_subreads = 14 if is_subjunc else 10
        p.top_scores = 3
        p.max_vote_simples = 64 if is_paired else 3
        p.max_vote_combinations = 3
        p.max_vote_number_cutoff = 2
        p.multi_best_reads = 3
        p.minimum_subread_for_first_read = 1 if is_subjunc else 3
        p.minimum_subread_for_second_read = 1
        p.max_indel_length = 5
        p.minimum_pair_distance = 50
        p.maximum_pair_distance = 600
        p.is_first_read_reversed = 0
        p.is_second_read_reversed = 1
        p.do_breakpoint_detection = 1 if is_subjunc else 0
        p.do_big_margin_filtering_for_junctions = 1 if is_subjunc else 0
        p.big_margin_record_size = 9
        p.maximum_intron_length = 500000
        p.limited_tree_scan = 0
        p.check_donor_at_junctions = 1
        p.more_accurate_fusions = 0       # ... && do_fusion_detection, core-interface-subjunc.c:671 / -aligner.c:648
        p.max_insertion_at_junctions = 0
        p.prefer_donor_receptor_junctions = 1
        return p


CORE_IS_GAPPED_READ = 64  # reference src/core.h:261


def used_slots_equal(a: np.ndarray, b: np.ndarray, ignore_gapped_flag: bool = False):
    """Byte-compare rule of SURVEY.md appendix A: slots with selected_votes > 0 must match in all
    bytes; for unused slots only ``selected_votes == 0`` is defined.  Returns (ok, first_bad_index).

    ``ignore_gapped_flag`` masks CORE_IS_GAPPED_READ in ``result_flags``: that bit is OR-ed in by
    the host-side ``find_new_indels`` / ``find_new_junctions`` (``mark_gapped_read``, reference
    src/core.h:61, src/core-indel.c:1979-2220) which run after the voting step inside ``do_voting``;
    bigtables dumped from the reference carry it, the voting step's own output does not."""
    a = a.reshape(-1).copy()
    b = b.reshape(-1).copy()
    if ignore_gapped_flag:
        a["result_flags"] &= ~np.int16(CORE_IS_GAPPED_READ)
        b["result_flags"] &= ~np.int16(CORE_IS_GAPPED_READ)
    if a.shape != b.shape:
        return False, -1
    used_a = a["selected_votes"] > 0
    used_b = b["selected_votes"] > 0
    bad = used_a != used_b
    raw_a = a.view(np.uint8).reshape(-1, MAPPING_RESULT.itemsize)
    raw_b = b.view(np.uint8).reshape(-1, MAPPING_RESULT.itemsize)
    # padding bytes (15, 66, 67) carry no information in either implementation
    keep = np.ones(MAPPING_RESULT.itemsize, dtype=bool)
    keep[[15, 66, 67]] = False
    diff = (raw_a[:, keep] != raw_b[:, keep]).any(axis=1) & used_a
    bad |= diff
    idx = np.flatnonzero(bad)
    return idx.size == 0, (int(idx[0]) if idx.size else -1)
