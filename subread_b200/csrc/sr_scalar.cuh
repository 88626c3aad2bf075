// sr_scalar.cuh -- the exact slow path: one thread replays one fragment with the reference's full
// 30 x 24 vote table held in a per-thread scratch area in global memory.
//
// The warp kernel (sr_warp.cuh) keeps at most 32 vote slots per table in shared memory; a fragment
// that needs more is queued and finished here, so the result is exact for any input (repeat-rich
// reads can fill up to 720 slots).  The same code also serves as an on-device cross-check
// (srcuda_set_force_slow_path).  It is written as host+device code so that tests can compile it
// with g++ and compare it with the oracle on a CPU-only box; the product only runs it on the GPU.
//
// Follows: do_voting core.c:3106-3235, gehash_go_X sorted-hashtable.c:937-1123,
// process_voting_junction_PE_topK core-junction.c:2199-2530.
#pragma once
#include "sr_common.cuh"

namespace sr {

struct ScalarTable {  // gene_vote_t (subread.h:324-354), live fields only
	int16_t max_vote;
	uint16_t items[kVoteRows];
	uint32_t pos[kVoteRows][kVoteSpace];
	int16_t votes[kVoteRows][kVoteSpace];
	int16_t last_cluster[kVoteRows][kVoteSpace];
	int16_t cov_start[kVoteRows][kVoteSpace];
	int16_t cov_end[kVoteRows][kVoteSpace];
	int8_t cursor[kVoteRows][kVoteSpace];
	int8_t toli[kVoteRows][kVoteSpace];
	uint8_t marked[kVoteRows][kVoteSpace];
	int16_t rec[kVoteRows][kVoteSpace][kRecLen];
};

struct ScalarSimple {  // simple_mapping_t core.h:414-424
	uint32_t pos;
	int16_t votes;
	int8_t is_vote_item;
	int8_t i, j;
};

struct ScalarCombo { int r1, r2; int score; };  // vote_combination_t core.h:426-431

struct ScalarScratch {
	ScalarTable vt[2];
	uint8_t text[2][kMaxReadLen + 6];
	uint32_t shift_locs[kVoteRows * kVoteSpace];
	ScalarSimple simple[2][kMaxSimples];
	ScalarCombo comb[kMaxCombos + 1];
	ResultWords tmp[2][kMaxBest];
	JuncWords jtmp[2][kMaxBest];
};

SR_HD int sc_char(uint8_t c) { return (int)(int8_t)c; }

// reverse_read, input-files.c:1171-1186
SR_HD uint8_t sc_complement(uint8_t c)
{
	switch (c) {
	case 'A': return 'T';
	case 'T': return 'A';
	case 'U': return 'A';
	case 'C': return 'G';
	case 'G': return 'C';
	default: return 'N';
	}
}

SR_HD void sc_reverse(uint8_t *s, int len)
{
	int i;
	for (i = 0; i < len / 2; i++) {
		uint8_t tail = s[len - 1 - i];
		s[len - 1 - i] = sc_complement(s[i]);
		s[i] = sc_complement(tail);
	}
	if (i * 2 == len - 1) s[i] = sc_complement(s[i]);
}

SR_HD uint32_t sc_key(const uint8_t *s)  // genekey2int input-files.c:1232
{
	uint32_t k = 0;
	for (int sh = 30; sh >= 0; sh -= 2) k |= (uint32_t)base_code(sc_char(*s++)) << sh;
	return k;
}

SR_HD void sc_table_init(ScalarTable *v)  // init_gene_vote gene-algorithms.h:42
{
	for (int i = 0; i < kVoteRows; i++) v->items[i] = 0;
	v->max_vote = 0;
}

SR_HD int sc_abs(int x) { return x < 0 ? -x : x; }

// gehash_go_X, sorted-hashtable.c:937-1123 (the table holds one strand at a time, so the mask
// comparison of :1019 is always true and is dropped).
SR_HD void sc_go_X(const DevIndex &ix, uint32_t raw_key, int offset, ScalarTable *vote, int tol, int subread_no,
                   uint32_t low_border, uint32_t high_border, int round, uint32_t *shift_locs, uint32_t *shift_n)
{
	uint32_t nb = ix.buckets_number;
#if defined(__CUDA_ARCH__)
	// the exact filters of the fast tiers (no false negatives) spare the bucket of most absent keys
	const uint32_t q = (uint32_t)__umul64hi((uint64_t)raw_key, ix.buckets_magic);
	int16_t key = (int16_t)q;
	uint32_t b = raw_key - q * nb;
	if (ix.direct) {
		uint32_t u;
		if (direct_bits(ix, raw_key, &u) == 0u) return;
	} else if (ix.bloom) {
		const uint32_t m = bloom_bits((int)key);
		if (((uint32_t)__ldg(ix.bloom + b) & m) != m) return;
	}
#else
	int16_t key = (int16_t)(raw_key / nb);
	uint32_t b = raw_key % nb;
#endif
	uint32_t base = ix.bucket_start[b];
	int items = (int)(ix.bucket_start[b + 1] - base);
	if (!items) return;
	const int16_t *keys = ix.keys + base;
	const uint32_t *vals = ix.vals + base;

	int imin = 0, imax = items - 1, at;
	for (;;) {
		at = (imin + imax) / 2;
		int16_t k = keys[at];
		if (k > key) imax = at - 1;
		else if (k < key) imin = at + 1;
		else break;
		if (imax < imin) return;
	}
	int sub_p1 = subread_no + 1, of_p16 = offset + 16, start_at = at, scan_back = 0;
	int ii_end = kSeg;
	if (tol > 5) ii_end = (tol % kSeg) ? (tol - tol % kSeg + kSeg) : tol;

	for (;;) {
		uint32_t kv = vals[at] - (uint32_t)offset;
		uint32_t home = vote_row(kv);
		int home_items = vote->items[home];
		bool found = false;
		for (int iix = 0; iix <= ii_end && !found; iix = iix > 0 ? -iix : (-iix + kSeg)) {
			uint32_t row = iix ? vote_row(kv + (uint32_t)iix) : home;
			int n = vote->items[row];
			for (int i = 0; i < n; i++) {
				int d = (int)(kv - vote->pos[row][i]);
				int t = (round > 0 && vote->marked[row][i]) ? 0 : tol;
				if (d < -t || d > t) continue;
				int toli = vote->toli[row][i];
				if (round == 0 && toli > 0 && d == 0 && !vote->marked[row][i]) {
					vote->marked[row][i] = 1;
					shift_locs[(*shift_n)++] = vote->pos[row][i];
				}
				if (sub_p1 == vote->last_cluster[row][i] && toli > 0) {
					int move = toli >= 3 ? vote->rec[row][i][toli - 1] : 0;
					int nd = move;
					move -= vote->rec[row][i][toli + 2];
					nd -= d;
					if (sc_abs(move) > sc_abs(nd)) {
						toli -= 3;
						vote->toli[row][i] = (int8_t)toli;
						vote->last_cluster[row][i]--;
						vote->votes[row][i]--;
					}
				}
				if (sub_p1 <= vote->last_cluster[row][i]) continue;
				int16_t nv = (int16_t)(vote->votes[row][i] + 1);
				vote->votes[row][i] = nv;
				if (of_p16 > vote->cov_end[row][i]) vote->cov_end[row][i] = (int16_t)of_p16;
				if (d == vote->cursor[row][i]) {
					vote->rec[row][i][toli + 1] = (int16_t)sub_p1;
				} else {
					toli += 3;
					if (toli < kMaxSections * 3) {
						vote->toli[row][i] = (int8_t)toli;
						vote->rec[row][i][toli] = (int16_t)sub_p1;
						vote->rec[row][i][toli + 1] = (int16_t)sub_p1;
						vote->rec[row][i][toli + 2] = (int16_t)d;
						if (toli < kMaxSections * 3 - 3) vote->rec[row][i][toli + 3] = 0;
					}
					vote->cursor[row][i] = (int8_t)d;
				}
				vote->last_cluster[row][i] = (int16_t)sub_p1;
				if (vote->max_vote < nv) vote->max_vote = nv;
				found = true;
				break;
			}
		}
		if (!found && kv >= low_border && kv <= high_border && home_items < kVoteSpace) {
			int j = home_items;
			vote->items[home]++;
			vote->pos[home][j] = kv;
			vote->votes[home][j] = 1;
			vote->toli[home][j] = 0;
			vote->marked[home][j] = 0;
			if (round > 0)
				for (uint32_t k = 0; k < *shift_n; k++)
					if (kv >= shift_locs[k] - (uint32_t)tol && kv <= shift_locs[k] + (uint32_t)tol) {
						vote->marked[home][j] = 1;
						break;
					}
			vote->rec[home][j][0] = vote->rec[home][j][1] = (int16_t)sub_p1;
			vote->rec[home][j][2] = 0;
			vote->rec[home][j][3] = 0;
			vote->cursor[home][j] = 0;
			vote->cov_start[home][j] = (int16_t)offset;
			vote->cov_end[home][j] = (int16_t)of_p16;
			vote->last_cluster[home][j] = (int16_t)sub_p1;
			if (vote->max_vote == 0) vote->max_vote = 1;
		}
		if (!scan_back) {
			at++;
			if (at == items || keys[at] != key) { scan_back = 1; at = start_at; }
		}
		if (scan_back) {
			at--;
			if (at < 0 || keys[at] != key) break;
		}
	}
}


SR_HD int sc_match_forward(const DevIndex &ix, const uint8_t *text, uint32_t pos, int len);

// ---- subjunc: minor half + donor site (core-junction.c:1073-1334, donor_score :3675-3832) --------------
SR_HD int sc_genome_char(const DevIndex &ix, uint32_t p)  // gvindex_get, gene-value-index.c:96-107
{
	uint32_t q = p - ix.start_base_offset;
	if ((q >> 2) >= ix.values_bytes - 1) return 'N';
	int code = (ix.genome[q >> 4] >> (2 * (q & 15))) & 3;
	return code == 0 ? 'A' : (code == 1 ? 'G' : (code == 2 ? 'C' : 'T'));
}

SR_HD bool sc_donor_chars(int a, int b) { return (a == 'G' && b == 'T') || (a == 'A' && b == 'G') || (a == 'A' && b == 'C') || (a == 'C' && b == 'T'); }

SR_HD int sc_donor_pair(int a0, int a1, int b0, int b1)  // paired_chars_part_core(.., 0)
{
	bool a_gt = a0 == 'G' && a1 == 'T', a_ag = a0 == 'A' && a1 == 'G', a_ct = a0 == 'C' && a1 == 'T', a_ac = a0 == 'A' && a1 == 'C';
	bool b_gt = b0 == 'G' && b1 == 'T', b_ag = b0 == 'A' && b1 == 'G', b_ct = b0 == 'C' && b1 == 'T', b_ac = b0 == 'A' && b1 == 'C';
	if ((a_gt && b_ag) || (a_ag && b_gt) || (a_ct && b_ac) || (a_ac && b_ct)) return (a_ct || a_gt) ? 1 : 0;
	return 0;
}

// The read as donor_score sees it: TextT::match(ix, off, pos, len) = bases of read[off, off + len) equal to the genome at pos
// (match_chro).  ScByteText is the scalar path's character buffer; the warp kernel passes its 2-bit packed array.
struct ScByteText {
	const uint8_t *text;
	SR_HD int match(const DevIndex &ix, int off, uint32_t pos, int len) const { return sc_match_forward(ix, text + off, pos, len); }
};

template <class TextT>
SR_HD int sc_donor_score_t(const DevIndex &ix, const VoteConfig &cfg, uint32_t left_pos, uint32_t right_pos, int left_indel, int right_indel,
                           int normally_arranged, int guess_start, int guess_end, const TextT &text, int rlen, int *split_point,
                           int *gt_ag_strand, int *donor_found)
{
	const int W = kConfirmWindow;
	int most_likely = (guess_start + guess_end) / 2;
	int sel_split = -1, sel_strand = -1, best = -111111;
	int dl0 = 0, dl1 = 0, dr0 = 0, dr1 = 0;
	int n_points = guess_end - guess_start;
	for (int k = 0; k < n_points; k++) {
		int sp = ((k % 2) ? -((k + 1) / 2) : ((1 + k) / 2)) + most_likely;
		int donor_ok = 0;
		if (sp > rlen - W) continue;
		if (sp < W) continue;
		if (cfg.prefer_donor) {
			if (normally_arranged) {
				uint32_t a = left_pos + (uint32_t)sp + (uint32_t)left_indel;
				dl0 = sc_genome_char(ix, a); dl1 = sc_genome_char(ix, a + 1);
				if (sc_donor_chars(dl0, dl1)) {
					uint32_t b = right_pos + (uint32_t)sp + (uint32_t)right_indel - 2;
					dr0 = sc_genome_char(ix, b); dr1 = sc_genome_char(ix, b + 1);
					if (sc_donor_chars(dr0, dr1)) donor_ok = sc_donor_pair(dl0, dl1, dr0, dr1);
				}
			} else {
				uint32_t a = right_pos + (uint32_t)sp + (uint32_t)left_indel, b = left_pos + (uint32_t)sp + (uint32_t)right_indel - 2;
				dl0 = sc_genome_char(ix, a); dl1 = sc_genome_char(ix, a + 1);
				dr0 = sc_genome_char(ix, b); dr1 = sc_genome_char(ix, b + 1);
				donor_ok = (sc_donor_chars(dl0, dl1) && sc_donor_chars(dr0, dr1)) ? sc_donor_pair(dl0, dl1, dr0, dr1) : 0;
			}
		}
		if (!(donor_ok || !cfg.check_donor)) continue;
		if (normally_arranged) {
			int lm = text.match(ix, sp - W, left_pos + (uint32_t)sp - W + (uint32_t)left_indel, W);
			if (lm > W - 2) {
				int rm = text.match(ix, sp, right_pos + (uint32_t)sp + (uint32_t)right_indel, W);
				if (rm >= 2 * W - lm - cfg.mismatch_allowed) {
					int lnm = text.match(ix, sp, left_pos + (uint32_t)sp + (uint32_t)left_indel, W);
					int rnm = text.match(ix, sp - W, right_pos + (uint32_t)sp + (uint32_t)right_indel - W, W);
					if (lnm <= W - 5 && rnm <= W - 5) {
						int score = 100 * (donor_ok * 3000 + lm + rm - lnm - rnm);
						if (score > best) { sel_strand = (dl0 == 'G' || dr1 == 'G') ? 1 : 0; sel_split = sp; best = score; }
					}
				}
			}
		} else {
			int rm = text.match(ix, sp - W, right_pos + (uint32_t)right_indel + (uint32_t)sp - W, W);
			int lm = text.match(ix, sp, left_pos + (uint32_t)sp + (uint32_t)left_indel, W);
			int rnm = text.match(ix, sp, right_pos + (uint32_t)sp + (uint32_t)right_indel, W);
			int lnm = text.match(ix, sp - W, left_pos + (uint32_t)left_indel + (uint32_t)sp - W, W);
			if (lm + rm >= 2 * W - cfg.mismatch_allowed && lnm <= W - 5 && rnm <= W - 5) {
				int score = 100 * (donor_ok * 3000 + lm + rm - lnm - rnm);
				if (score > best) { sel_strand = (dl0 == 'G' || dr1 == 'G') ? 1 : 0; sel_split = sp; best = score; }
			}
		}
	}
	if (best > 0) {
		*split_point = sel_split;
		*donor_found = best >= 290000;
		*gt_ag_strand = sel_strand;
		return (1 + best) / 100;
	}
	return 0;
}

SR_HD int sc_donor_score(const DevIndex &ix, const VoteConfig &cfg, uint32_t left_pos, uint32_t right_pos, int left_indel, int right_indel,
                         int normally_arranged, int guess_start, int guess_end, const uint8_t *text, int rlen, int *split_point,
                         int *gt_ag_strand, int *donor_found)
{
	return sc_donor_score_t(ix, cfg, left_pos, right_pos, left_indel, right_indel, normally_arranged, guess_start, guess_end, ScByteText{text},
	                        rlen, split_point, gt_ag_strand, donor_found);
}

SR_HD void sc_big_margin_insert(const VoteConfig &cfg, uint16_t *rec, int votes, int start, int end, int rlen, bool negative)
{   // insert_big_margin_record, core-junction.c:789-813 (votes is an unsigned char there)
	const int n = cfg.big_margin_record_size;
	if (n < 3) return;
	const uint16_t v8 = (uint16_t)(uint8_t)votes;
	const uint16_t s2 = (uint16_t)(negative ? rlen - end : start), e2 = (uint16_t)(negative ? rlen - start : end);
	int x;
	for (x = 0; x < n / 3; x++)
		if (v8 >= rec[x * 3]) break;
	if (x < n / 3) {
		for (int y = n - 4; y >= x * 3; y--) rec[y + 3] = rec[y];
		rec[x * 3] = v8; rec[x * 3 + 1] = s2; rec[x * 3 + 2] = e2;
	}
}

// The do_breakpoint_detection part of copy_vote_to_alignment_res for one emitted slot.
SR_HD void sc_minor_half(const DevIndex &ix, const VoteConfig &cfg, ResultWords *r, JuncWords *jr, const ScalarTable *v, int vi, int vj,
                         int rlen, const uint8_t *text)
{
	const int sel_votes = (int16_t)(r->w[2] & 0xffffu);
	for (int i = 0; i < kVoteRows; i++)
		for (int j = 0; j < v->items[i]; j++) {
			if (i == vi && j == vj) continue;
			if (sel_votes < v->votes[i][j]) continue;
			const uint32_t pm = v->pos[vi][vj], pn = v->pos[i][j];
			const long long dist = (long long)pm - (long long)pn;
			const int cs_m = v->cov_start[vi][vj], ce_m = v->cov_end[vi][vj], cs_n = v->cov_start[i][j], ce_n = v->cov_end[i][j];
			if (sc_abs((int)dist) > cfg.maximum_intron_length) continue;  // test_junction_minor :889 (abs() truncates to int)
			if (cs_m == cs_n) continue;
			if (ce_m == ce_n) continue;
			if (cs_m > cs_n) { if (pm < pn) continue; } else { if (pm > pn) continue; }
			// is_better_inner :975
			const int16_t jr_votes = (int16_t)(jr->w[0] >> 16);
			const int jcov = (int)(uint16_t)(jr->w[3] >> 16) - (int)(uint16_t)(jr->w[3] & 0xffffu);
			uint32_t oi = pm - jr->w[2]; if (oi > 0x7fffffffu) oi = (0xffffffffu - oi) + 1;
			uint32_t ni = pm - pn; if (ni > 0x7fffffffu) ni = (0xffffffffu - ni) + 1;
			const int mv = v->votes[i][j], mcov = ce_n - cs_n;
			const bool better = mv > jr_votes || (mv == jr_votes && mcov > jcov) || (mv == jr_votes && mcov == jcov && (int)ni < (int)oi);
			if (!better) continue;
			const int overlapped = cs_m > cs_n ? ce_n - cs_m : ce_m - cs_n;
			if (overlapped > 14) continue;
			if (sc_abs((int)dist) < 6) continue;
			const int guess_start = cs_m > cs_n ? ce_n - 8 : ce_m - 8;
			const int guess_end = cs_m < cs_n ? cs_n + 8 : cs_m + 8;
			const int normally_arranged = 1 != (cs_m > cs_n) + (pm > pn);
			int left_indel = 0, major_indels = 0, minor_indel_offset = 0;
			if (rlen > kLongRead) {
				for (int k = 0; k < kMaxSections; k++) { if (!v->rec[vi][vj][k * 3]) break; major_indels += v->rec[vi][vj][k * 3 + 2]; }
				for (int k = 0; k < kMaxSections; k++) { if (!v->rec[i][j][k * 3]) break; minor_indel_offset += v->rec[i][j][k * 3 + 2]; }
				left_indel = pm < pn ? major_indels : minor_indel_offset;
			}
			int split = 0, gt_ag = 0, donor_found = 0;
			const int replace = sc_donor_score(ix, cfg, pm < pn ? pm : pn, pm > pn ? pm : pn, left_indel, 0, normally_arranged,
			                                   guess_start > 0 ? guess_start : 0, guess_end < rlen ? guess_end : rlen, text, rlen, &split, &gt_ag, &donor_found);
			if (!replace) continue;
			jr->w[0] = (uint32_t)(uint16_t)split | ((uint32_t)(uint16_t)mv << 16);
			jr->w[1] = (uint32_t)(uint8_t)((minor_indel_offset & 0xf) | ((major_indels & 0xf) << 4)) | (0u << 8) |
			           ((uint32_t)(uint8_t)(!normally_arranged) << 16) | ((uint32_t)(uint8_t)normally_arranged << 24);
			jr->w[2] = pn;
			jr->w[3] = (uint32_t)(uint16_t)cs_n | ((uint32_t)(uint16_t)ce_n << 16);
			uint32_t flags = r->w[1] & 0xffffu;
			flags &= ~0x3u;
			if (!donor_found || gt_ag > 2) flags |= 3u;
			else flags = gt_ag ? (flags | 1u) : (flags & ~1u);
			flags &= ~4u;
			r->w[1] = (r->w[1] & 0xffff0000u) | flags;
		}
}

SR_HD void sc_update_top(int top_scores, int *top, int v)  // update_top_three core-junction.c:908
{
	if (v > top[top_scores - 1])
		for (int x = 0; x < top_scores; x++) {
			if (v > top[x]) {
				for (int y = top_scores - 1; y > x; y--) top[y] = top[y - 1];
				top[x] = v;
				break;
			} else if (v == top[x]) break;
		}
}

// copy_vote_to_alignment_res (core-junction.c:1058-1072) into the 17-word image of a zeroed
// mapping_result_t, including indel_recorder_copy (sorted-hashtable.c:1144-1164).
SR_HD void sc_slot_to_result(ResultWords *r, const ScalarTable *v, int i, int j, int used_subreads)
{
	int16_t rec[22];
	for (int k = 0; k < 22; k++) rec[k] = 0;
	int16_t all = 0;
	int n = 0;
	const int16_t *src = v->rec[i][j];
	while (src[n] && n < kRecLen - 2) {
		rec[n] = src[n]; n++;
		rec[n] = src[n]; n++;
		rec[n] = src[n]; all = rec[n]; n++;
	}
	for (int k = 0; k < kResultWords; k++) r->w[k] = 0;
	r->w[0] = v->pos[i][j];
	r->w[1] = 0;  // strand flag is set by the caller (the table does not store it)
	r->w[2] = (uint32_t)(uint16_t)v->votes[i][j] | ((uint32_t)(uint16_t)used_subreads << 16);
	r->w[3] = ((uint32_t)(uint8_t)(int8_t)all) << 8;
	for (int k = 0; k < 11; k++) r->w[4 + k] = (uint32_t)(uint16_t)rec[2 * k] | ((uint32_t)(uint16_t)rec[2 * k + 1] << 16);
	r->w[15] = (uint32_t)(uint16_t)v->cov_start[i][j] | ((uint32_t)(uint16_t)v->cov_end[i][j] << 16);
}

// process_voting_junction_PE_topK, core-junction.c:2199-2530, DNA mode.
SR_HD void sc_topk(const DevIndex &ix, const VoteConfig &cfg, int strand, ScalarTable *const vt[2], const int rlen[2],
                   const int all_subreads[2], ResultWords *res, JuncWords *jres, uint16_t *big_margin, ScalarScratch *sc)
{
	const bool junc = cfg.do_junctions && jres != nullptr;
	const int mates = 1 + cfg.is_paired, mb = cfg.multi_best_reads;
	int top[2][kMaxTop];
	int n_simple[2] = {0, 0};

	for (int m = 0; m < mates; m++) {
		for (int k = 0; k < cfg.top_scores; k++) top[m][k] = 0;
		for (int i = 0; i < kVoteRows; i++)
			for (int j = 0; j < vt[m]->items[i]; j++) sc_update_top(cfg.top_scores, top[m], vt[m]->votes[i][j]);
		for (int k = 0; k < mb; k++)
			if (result_votes(res[m * mb + k]) > 0) sc_update_top(cfg.top_scores, top[m], result_votes(res[m * mb + k]));
	}
	for (int m = 0; m < mates; m++) {
		int n = 0;
		ScalarSimple *sp = sc->simple[m];
		for (int lv = 0; lv < cfg.top_scores; lv++) {
			if (n >= cfg.max_vote_simples) break;
			int want = top[m][lv];
			if (want < 1 || top[m][0] - want > cfg.max_vote_number_cutoff) break;
			for (int i = 0; i < kVoteRows && n < cfg.max_vote_simples; i++)
				for (int j = 0; j < vt[m]->items[i] && n < cfg.max_vote_simples; j++) {
					int vv = vt[m]->votes[i][j];
					if (cfg.big_margin_filtering && big_margin && lv == 0 && vv >= top[m][cfg.top_scores - 1])  // core-junction.c:2281
						sc_big_margin_insert(cfg, big_margin + m * kBigMarginU16, vv, vt[m]->cov_start[i][j], vt[m]->cov_end[i][j], rlen[m], strand != 0);
					if (vv == want && vv >= cfg.min_second) {
						sp[n].is_vote_item = 1; sp[n].i = (int8_t)i; sp[n].j = (int8_t)j;
						sp[n].pos = vt[m]->pos[i][j]; sp[n].votes = (int16_t)vv;
						n++;
					}
				}
			for (int k = 0; k < mb && n < cfg.max_vote_simples; k++)
				if (result_votes(res[m * mb + k]) == want) {
					sp[n].is_vote_item = 0; sp[n].i = (int8_t)k; sp[n].j = 0;
					sp[n].pos = res[m * mb + k].w[0]; sp[n].votes = result_votes(res[m * mb + k]);
					n++;
				}
		}
		n_simple[m] = n;
	}

	int n_comb = 0;
	ScalarCombo *comb = sc->comb;
	if (cfg.is_paired) {
		for (int i = 0; i < n_simple[0]; i++) {
			int c1, p1, re1 = locate_position(ix.chr_offsets, ix.total_offsets, ix.padding, sc->simple[0][i].pos, &c1, &p1);
			for (int j = 0; j < n_simple[1]; j++) {
				int v1 = sc->simple[0][i].votes, v2 = sc->simple[1][j].votes;
				if ((v1 > v2 ? v1 : v2) < cfg.min_first) continue;
				int c2, p2, re2 = locate_position(ix.chr_offsets, ix.total_offsets, ix.padding, sc->simple[1][j].pos, &c2, &p2);
				int is_pe;
				int weight = pair_weight(re1, c1, p1, re2, c2, p2, rlen[0], rlen[1], cfg.min_pair_dist, cfg.max_pair_dist, &is_pe);
				if (!is_pe && (v1 < v2 ? v1 : v2) < cfg.min_first) continue;
				int score = (v1 + v2) * weight;
				int t;
				for (t = 0; t < n_comb; t++)
					if (comb[t].score < score) break;
				if (t < cfg.max_vote_combinations) {
					int last = n_comb < cfg.max_vote_combinations - 1 ? n_comb : cfg.max_vote_combinations - 1;
					for (int mv = last; mv > t; mv--) comb[mv] = comb[mv - 1];
					comb[t].r1 = i; comb[t].r2 = j; comb[t].score = score;
					if (n_comb < cfg.max_vote_combinations) n_comb++;
				}
			}
		}
	}

	int cursor[2] = {0, 0};
	const uint32_t strand_flag = strand ? (uint32_t)kCoreNegStrand : 0u;
	if (n_comb > 0) {
		// merge_sort -> basic_sort_run for <= 11 items (core.c:4716-4749): unstable selection sort
		for (int i = 0; i < n_comb - 1; i++) {
			int mn = i;
			for (int j = i + 1; j < n_comb; j++)
				if (comb[mn].score - comb[j].score > 0) mn = j;
			if (i != mn) { ScalarCombo t = comb[i]; comb[i] = comb[mn]; comb[mn] = t; }
		}
		for (int m = 0; m < mates; m++)
			for (int c = n_comb - 1; c >= 0 && cursor[m] < mb; c--) {
				const ScalarSimple *loc = &sc->simple[m][m ? comb[c].r2 : comb[c].r1];
				bool exist = false;
				for (int j = 0; j < cursor[m]; j++)
					if (sc->tmp[m][j].w[0] == loc->pos) { exist = true; break; }
				if (exist) continue;
				if (junc) for (int w = 0; w < 4; w++) sc->jtmp[m][cursor[m]].w[w] = 0;
				if (loc->is_vote_item) {
					sc_slot_to_result(&sc->tmp[m][cursor[m]], vt[m], loc->i, loc->j, all_subreads[m]);
					sc->tmp[m][cursor[m]].w[1] = strand_flag;
					if (junc) sc_minor_half(ix, cfg, &sc->tmp[m][cursor[m]], &sc->jtmp[m][cursor[m]], vt[m], loc->i, loc->j, rlen[m], sc->text[m]);
				} else {
					sc->tmp[m][cursor[m]] = res[m * mb + loc->i];
					if (junc) sc->jtmp[m][cursor[m]] = jres[m * mb + loc->i];
				}
				cursor[m]++;
			}
	} else {
		if (n_simple[0] == 0) res[0].w[3] &= ~0xffu;
		if (cfg.is_paired && n_simple[1] == 0) res[mb].w[3] &= ~0xffu;
		if (n_simple[0] > 0 || n_simple[1] > 0)
			for (int m = 0; m < mates; m++)
				for (int s = 0; s < n_simple[m] && cursor[m] < mb; s++) {
					const ScalarSimple *loc = &sc->simple[m][s];
					if (loc->votes < cfg.min_first) continue;
					bool exist = false;
					for (int j = 0; j < cursor[m]; j++)
						if (sc->tmp[m][j].w[0] == loc->pos) { exist = true; break; }
					if (exist) continue;
					if (junc) for (int w = 0; w < 4; w++) sc->jtmp[m][cursor[m]].w[w] = 0;
					if (loc->is_vote_item) {
						sc_slot_to_result(&sc->tmp[m][cursor[m]], vt[m], loc->i, loc->j, all_subreads[m]);
						sc->tmp[m][cursor[m]].w[1] = strand_flag;
						if (junc) sc_minor_half(ix, cfg, &sc->tmp[m][cursor[m]], &sc->jtmp[m][cursor[m]], vt[m], loc->i, loc->j, rlen[m], sc->text[m]);
					} else {
						sc->tmp[m][cursor[m]] = res[m * mb + loc->i];
						if (junc) sc->jtmp[m][cursor[m]] = jres[m * mb + loc->i];
					}
					cursor[m]++;
				}
	}
	for (int m = 0; m < mates; m++)
		for (int k = 0; k < mb; k++) {
			if (k < cursor[m]) res[m * mb + k] = sc->tmp[m][k];
			else res[m * mb + k].w[2] &= 0xffff0000u;  // selected_votes = 0 only (core-junction.c:2517)
			if (junc) {
				if (k < cursor[m]) jres[m * mb + k] = sc->jtmp[m][k];
				else jres[m * mb + k].w[0] &= 0x0000ffffu;  // minor_votes = 0 (core-junction.c:2524)
			}
		}
}

// Number of read bases equal to the genome (match_chro, gene-value-index.c:856-959, forward
// branch; the text has been through reverse_read at least once, so it holds only A,C,G,T,N).
SR_HD int sc_match_forward(const DevIndex &ix, const uint8_t *text, uint32_t pos, int len)
{
	if ((uint32_t)(pos + (uint32_t)len) >= ix.length + ix.start_point) return 0;
	if (pos > 0xffff0000u) return 0;
	uint32_t q = pos - ix.start_base_offset;
	if ((q >> 2) >= ix.values_bytes) return 0;
	int ret = 0;
	for (int i = 0; i < len; i++, q++) {
		int g = (ix.genome[q >> 4] >> (2 * (q & 15))) & 3;
		int c = text[i];
		int code = c == 'A' ? 0 : (c == 'G' ? 1 : (c == 'C' ? 2 : (c == 0 ? -1 : 3)));
		ret += g == code;
	}
	return ret;
}

// One fragment through do_voting's body (core.c:3106-3235).  res is read-modify-write
// [mates * multi_best_reads]; verify (optional) gets one count per result slot.
SR_HD void scalar_vote_fragment(const DevIndex &ix, const VoteConfig &cfg, const uint8_t *raw0, int len0,
                                const uint8_t *raw1, int len1, ResultWords *res, JuncWords *jres, uint16_t *big_margin, uint16_t *verify,
                                ScalarScratch *sc)
{
	const int mates = 1 + cfg.is_paired, mb = cfg.multi_best_reads, gap = ix.index_gap;
	ScalarTable *vt[2] = {&sc->vt[0], &sc->vt[1]};
	const uint8_t *raw[2] = {raw0, raw1};
	int rlen[2] = {len0, len1};
	int all_subreads[2] = {0, 0};

	for (int m = 0; m < mates; m++) {  // fetch_next_read_pair core.c:1187-1199
		for (int i = 0; i < rlen[m]; i++) sc->text[m][i] = raw[m][i];
		for (int i = rlen[m]; i < rlen[m] + 6; i++) sc->text[m][i] = 0;
		if (m == 0 ? cfg.first_reversed : cfg.second_reversed) sc_reverse(sc->text[m], rlen[m]);
	}
	for (int strand = 0; strand < 2; strand++) {
		int applied = 0;
		for (int m = 0; m < mates; m++) {
			sc_table_init(vt[m]);
			if (rlen[m] < 16) continue;
			int step;
			applied = subread_schedule(rlen[m], gap, cfg.total_subreads, &step);
			all_subreads[m] = applied;
			uint32_t cur_high = cfg.high_border - (uint32_t)rlen[m];
			uint32_t shift_n = 0;
			for (int round = 0; round < 2; round++) {
				sc_table_init(vt[m]);
				for (int l = 0; l < applied * gap; l++) {
					int o = lookup_offset(l, gap, step);
					sc_go_X(ix, sc_key(sc->text[m] + o), o, vt[m], cfg.indel_tolerance, l / gap, cfg.low_border, cur_high, round,
					        sc->shift_locs, &shift_n);
				}
				if (shift_n == 0) break;
			}
		}
		if (cfg.is_paired) {
			sc_topk(ix, cfg, strand, vt, rlen, all_subreads, res, jres, big_margin, sc);
		} else if (vt[0]->max_vote >= cfg.min_first) {
			sc_topk(ix, cfg, strand, vt, rlen, all_subreads, res, jres, big_margin, sc);
		} else if (result_votes(res[0]) < 1) {  // core.c:3217-3225
			res[0].w[3] &= ~0xffu;
			int16_t used = (int16_t)(res[0].w[2] >> 16);
			if (used < applied) res[0].w[2] = (res[0].w[2] & 0xffffu) | ((uint32_t)(uint16_t)applied << 16);
		}
		if (strand == 0)
			for (int m = 0; m < mates; m++) sc_reverse(sc->text[m], rlen[m]);
	}
	if (verify)
		for (int m = 0; m < mates; m++)
			for (int k = 0; k < mb; k++) {
				const ResultWords &r = res[m * mb + k];
				if (result_votes(r) < 1) { verify[m * mb + k] = 0xffff; continue; }
				bool neg = (r.w[1] & kCoreNegStrand) != 0;
				if (!neg) sc_reverse(sc->text[m], rlen[m]);
				verify[m * mb + k] = (uint16_t)sc_match_forward(ix, sc->text[m], r.w[0], rlen[m]);
				if (!neg) sc_reverse(sc->text[m], rlen[m]);
			}
}

}  // namespace sr
