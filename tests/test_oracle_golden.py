"""The oracle (oracle/sr_oracle.c) against bigtables the unmodified reference produced
(tests/golden/make_golden.py): pins the checker itself.  CPU only."""
import pytest

from conftest import describe_mismatch, junc_tables_equal, load_case, load_case_junc
from subread_b200 import used_slots_equal

CASE_NAMES = ["se_gapped", "pe_gapped", "se_full", "pe_full", "se_I16", "pe_n14_ff"]


@pytest.mark.parametrize("name", CASE_NAMES)
def test_oracle_matches_reference_bigtable(name, golden_cases, golden_index_dir, oracle_mod):
    case = next(c for c in golden_cases if c["name"] == name)
    prefix, chunk, params, exp = load_case(case, golden_index_dir)
    idx = oracle_mod.OracleIndex(prefix)
    got, _, stats = oracle_mod.vote_chunk(idx, params, chunk)
    ok, bad = used_slots_equal(got, exp)
    assert ok, describe_mismatch(got, exp, bad)
    # every byte, not only the used slots (the reference memcpy's zeroed temporaries)
    assert got.tobytes() == exp.tobytes()
    assert stats.lookups > 0 and stats.algorithmic_bytes() > 0


def test_oracle_threads_agree(golden_cases, golden_index_dir, oracle_mod):
    case = next(c for c in golden_cases if c["name"] == "pe_gapped")
    prefix, chunk, params, exp = load_case(case, golden_index_dir)
    idx = oracle_mod.OracleIndex(prefix)
    got, _, _ = oracle_mod.vote_chunk(idx, params, chunk, threads=4)
    assert got.tobytes() == exp.tobytes()


def test_oracle_second_pass_keeps_the_best_location(golden_cases, golden_index_dir, oracle_mod):
    """Voting the same block a second time merges every table with its own earlier result
    (out_align is read-modify-write).  The best slot must not move and no new location may appear;
    lower slots may legitimately be lost, because old results and table slots at the same vote
    level share the max_vote_simples budget (core-junction.c:2256-2316) before deduplication."""
    case = next(c for c in golden_cases if c["name"] == "se_gapped")
    prefix, chunk, params, exp = load_case(case, golden_index_dir)
    idx = oracle_mod.OracleIndex(prefix)
    once, _, _ = oracle_mod.vote_chunk(idx, params, chunk)
    twice, _, _ = oracle_mod.vote_chunk(idx, params, chunk, out_align=once.copy())
    best1, best2 = once[:, :, 0], twice[:, :, 0]
    assert (best1["selected_votes"] == best2["selected_votes"]).all()
    used = best1["selected_votes"] > 0
    assert (best1["selected_position"][used] == best2["selected_position"][used]).all()
    for f in range(chunk.n_frags):
        before = {int(p) for p, v in zip(once[f, 0]["selected_position"], once[f, 0]["selected_votes"]) if v > 0}
        after = {int(p) for p, v in zip(twice[f, 0]["selected_position"], twice[f, 0]["selected_votes"]) if v > 0}
        assert after <= before


@pytest.mark.parametrize("name", ["sj_se", "sj_pe"])
def test_oracle_subjunc_matches_reference(name, golden_cases, golden_index_dir, oracle_mod):
    """subjunc (do_breakpoint_detection): alignment table, minor-half / donor results and big-margin records
    against the reference's ref_dump_subjunc on reads spanning planted GT..AG / CT..AC introns."""
    case = next(c for c in golden_cases if c["name"] == name)
    prefix, chunk, params, exp = load_case(case, golden_index_dir)
    exp_junc, exp_bm = load_case_junc(case, chunk, params)
    assert (exp_junc["minor_votes"] > 0).sum() > 20
    idx = oracle_mod.OracleIndex(prefix)
    js = {}
    got, _, _ = oracle_mod.vote_chunk(idx, params, chunk, junc_state=js)
    ok, bad = used_slots_equal(got, exp)
    assert ok, describe_mismatch(got, exp, bad)
    assert junc_tables_equal(got, js["junc"], js["big_margin"], exp_junc, exp_bm)
