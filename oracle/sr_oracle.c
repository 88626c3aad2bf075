/* oracle/sr_oracle.c -- TEST INFRASTRUCTURE ONLY (see sr_oracle.h).
 *
 * Plain-C restatement of the reference's voting step.  Every function names the reference
 * file:line it follows.  The structure deliberately mirrors the reference's sequential
 * semantics (30x24 vote table, first-accepting-match wins, visit order of equal-key runs)
 * because bit-exactness depends on them.
 *
 * Parity status: PINNED against oracle/_ref (tests/test_oracle_vs_ref.py) and committed
 * golden bigtables (tests/golden/).
 */
#include "sr_oracle.h"
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define VT_ROWS 30            /* GENE_VOTE_TABLE_SIZE  subread.h:217 */
#define VT_SPACE 24           /* GENE_VOTE_SPACE       subread.h:216 */
#define MAX_SECTIONS 7        /* MAX_INDEL_SECTIONS    subread.h:88  */
#define REC_LEN 21            /* MAX_INDEL_TOLERANCE*3 subread.h:231 */
#define SEG 5                 /* INDEL_SEGMENT_SIZE    sorted-hashtable.c:344 */
#define NEG_STRAND_MASK 2048  /* IS_NEGATIVE_STRAND    subread.h:112 */
#define CORE_NEG 8            /* CORE_IS_NEGATIVE_STRAND core.h:258 */
#define LONG_READ 160         /* EXON_LONG_READ_LENGTH subread.h:83 */
#define MAX_RLEN 1210         /* MAX_READ_LENGTH       subread.h:73 */
#define MAX_TOP 9

/* gene_vote_t, the fields the live path touches (subread.h:324-354). */
typedef struct {
	int16_t max_vote;
	uint16_t items[VT_ROWS];
	uint32_t pos[VT_ROWS][VT_SPACE];
	int32_t masks[VT_ROWS][VT_SPACE];
	int32_t marked_shift_indel[VT_ROWS][VT_SPACE];
	int16_t votes[VT_ROWS][VT_SPACE];
	int16_t last_subread_cluster[VT_ROWS][VT_SPACE];
	int16_t indel_recorder[VT_ROWS][VT_SPACE][REC_LEN];
	int8_t current_indel_cursor[VT_ROWS][VT_SPACE];
	int8_t toli[VT_ROWS][VT_SPACE];
	int16_t coverage_start[VT_ROWS][VT_SPACE];
	int16_t coverage_end[VT_ROWS][VT_SPACE];
} vote_table_t;

/* init_gene_vote, gene-algorithms.h:42 */
static void vote_init(vote_table_t *v)
{
	memset(v->items, 0, sizeof(v->items));
	v->max_vote = 0;
}

void or_default_params(or_params_t *p, int is_paired, int is_subjunc)
{
	memset(p, 0, sizeof(*p));
	/* init_global_context, core-indel.c:4399-4538 */
	p->total_subreads = 10;
	p->top_scores = 3;
	p->minimum_subread_for_first_read = 3;
	p->minimum_subread_for_second_read = 1;
	p->max_indel_length = 5;
	p->minimum_pair_distance = 50;
	p->maximum_pair_distance = 600;
	p->is_first_read_reversed = 0;
	p->is_second_read_reversed = 1;
	p->big_margin_record_size = 9;
	p->maximum_intron_length = 500000;
	p->check_donor_at_junctions = 1;
	p->prefer_donor_receptor_junctions = 1;
	/* load_global_context, core.c:4075-4084 */
	p->max_vote_combinations = 3;
	p->multi_best_reads = 3;
	p->max_vote_simples = is_paired ? 64 : 3;
	p->max_vote_number_cutoff = 2;
	if (is_subjunc) { /* parse_opts_subjunc, core-interface-subjunc.c:264-284 */
		p->total_subreads = 14;
		p->minimum_subread_for_first_read = 1;
		p->minimum_subread_for_second_read = 1;
		p->do_breakpoint_detection = 1;
		p->do_big_margin_filtering_for_junctions = 1;
		p->limited_tree_scan = 0;
	}
}

/* ------------------------------------------------------------------ index files */

static int read_exact(FILE *fp, void *dst, size_t n) { return fread(dst, 1, n, fp) == n ? 0 : -1; }

/* gehash_load (VER2 branch), sorted-hashtable.c:1424-1568. */
static int load_tab(or_index_t *idx, const char *path)
{
	FILE *fp = fopen(path, "rb");
	char magic[8];
	if (!fp) return 1;
	if (read_exact(fp, magic, 8) || memcmp(magic, "2subindx", 8)) { fclose(fp); return 2; }
	idx->index_gap = 0;
	idx->padding = 0;
	for (;;) {
		int16_t okey, olen;
		if (read_exact(fp, &okey, 2)) { fclose(fp); return 3; }
		if (!okey) break;
		if (read_exact(fp, &olen, 2)) { fclose(fp); return 3; }
		if (okey == 0x0101) { int16_t v; if (read_exact(fp, &v, 2)) { fclose(fp); return 3; } idx->index_gap = v; }
		else if (okey == 0x0102) { int16_t v; if (read_exact(fp, &v, 2)) { fclose(fp); return 3; } idx->padding = v; }
		else fseek(fp, olen, SEEK_CUR);
	}
	int64_t items;
	int32_t nb;
	if (read_exact(fp, &items, 8) || read_exact(fp, &nb, 4)) { fclose(fp); return 3; }
	idx->n_items = (uint64_t)items;
	idx->buckets_number = (uint32_t)nb;
	idx->bucket_start = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)nb + 1));
	idx->keys = (int16_t *)malloc(sizeof(int16_t) * (size_t)(items ? items : 1));
	idx->vals = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(items ? items : 1));
	uint64_t at = 0;
	for (int32_t b = 0; b < nb; b++) {
		int32_t hdr[2];
		if (read_exact(fp, hdr, 8)) { fclose(fp); return 4; }
		idx->bucket_start[b] = (uint32_t)at;
		if (at + (uint64_t)hdr[0] > idx->n_items) { fclose(fp); return 5; }
		if (hdr[0] > 0) {
			if (read_exact(fp, idx->keys + at, 2 * (size_t)hdr[0])) { fclose(fp); return 4; }
			if (read_exact(fp, idx->vals + at, 4 * (size_t)hdr[0])) { fclose(fp); return 4; }
		}
		at += (uint64_t)hdr[0];
	}
	idx->bucket_start[nb] = (uint32_t)at;
	fclose(fp);
	return at == idx->n_items ? 0 : 6;
}

/* gvindex_load, gene-value-index.c:190-228. */
static int load_array(or_index_t *idx, const char *path)
{
	FILE *fp = fopen(path, "rb");
	if (!fp) return 1;
	if (read_exact(fp, &idx->start_point, 4) || read_exact(fp, &idx->length, 4)) { fclose(fp); return 2; }
	idx->start_base_offset = idx->start_point - idx->start_point % 4;
	uint32_t useful_bytes = (idx->length + idx->start_point - idx->start_base_offset) >> 2;
	idx->values_bytes = useful_bytes + 1;
	idx->values = (uint8_t *)calloc(idx->values_bytes + 16, 1);
	size_t got = fread(idx->values, 1, idx->values_bytes, fp);
	fclose(fp);
	return got < useful_bytes ? 3 : 0;
}

/* load_offsets, gene-algorithms.c:1293-1372 (only the numeric column matters here). */
static int load_reads(or_index_t *idx, const char *path)
{
	FILE *fp = fopen(path, "r");
	char line[4096];
	int cap = 128, n = 0;
	if (!fp) return 1;
	idx->read_offsets = (uint32_t *)malloc(sizeof(uint32_t) * cap);
	while (fgets(line, sizeof(line), fp)) {
		if (strlen(line) < 2) continue;
		char *tab = strchr(line, '\t');
		if (!tab) continue;
		*tab = 0;
		if (n + 1 >= cap) { cap *= 2; idx->read_offsets = (uint32_t *)realloc(idx->read_offsets, sizeof(uint32_t) * cap); }
		idx->read_offsets[n++] = (uint32_t)atoll(line);
	}
	fclose(fp);
	idx->total_offsets = n;
	return 0;
}

int or_index_load(or_index_t *idx, const char *tab_path, const char *array_path, const char *reads_path)
{
	int r;
	memset(idx, 0, sizeof(*idx));
	if ((r = load_tab(idx, tab_path))) return 100 + r;
	if ((r = load_array(idx, array_path))) return 200 + r;
	if ((r = load_reads(idx, reads_path))) return 300 + r;
	return 0;
}

void or_index_free(or_index_t *idx)
{
	free(idx->bucket_start); free(idx->keys); free(idx->vals); free(idx->values); free(idx->read_offsets);
	memset(idx, 0, sizeof(*idx));
}

/* ------------------------------------------------------------------ read encoding */

/* base2int, subread.h:238 -- note the argument is a (signed) char in genekey2int. */
static inline int base_code(char c) { return c < 'G' ? (c == 'A' ? 0 : 2) : (c == 'G' ? 1 : 3); }

/* genekey2int, input-files.c:1232-1251 (base space): first base in bits 31..30. */
uint32_t or_genekey2int(const char *s)
{
	uint32_t key = 0;
	for (int sh = 30; sh >= 0; sh -= 2) key |= (uint32_t)base_code(*s++) << sh;
	return key;
}

/* The 256-entry table at input-files.c:1111 as a rule: A<->T, C<->G, U->A, everything else
 * (lowercase letters included: the second "TNGNNNC" group of the table sits at 0xC1.., not at
 * 'a'..) -> N.  Checked for every byte 1..127 against the reference in tests/test_oracle_vs_ref.py. */
static inline char complement_char(unsigned char c)
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

/* reverse_read, input-files.c:1171-1186 (base space branch). */
void or_reverse_read(char *s, int len)
{
	int i;
	for (i = 0; i < len / 2; i++) {
		unsigned char tail = (unsigned char)s[len - 1 - i];
		s[len - 1 - i] = complement_char((unsigned char)s[i]);
		s[i] = complement_char(tail);
	}
	if (i * 2 == len - 1) s[i] = complement_char((unsigned char)s[i]);
}

/* do_voting's subread schedule, core.c:3117-3130 and :3169-3171. */
int or_subread_schedule(int rlen, int gap, int total_subreads, int *applied_subreads, int *offsets, int max_out)
{
	int n = 0;
	*applied_subreads = 0;
	if (rlen < 16) return 0;
	int cr15 = (rlen - 15 - gap) * 65536;
	int step;
	if (rlen <= LONG_READ) {
		step = cr15 / (total_subreads - 1);
		if (step < gap * 65536) step = gap * 65536;
	} else {
		step = 6 * 65536;
		if (cr15 / step > 62) step = cr15 / 62;
	}
	int applied = 1 + cr15 / step;
	*applied_subreads = applied;
	for (int s = 0; s < applied; s++)
		for (int x = 0; x < gap; x++) {
			int o = (step * s) >> 16;
			if (gap > 1) o -= o % gap - x;
			if (n < max_out) offsets[n] = o;
			n++;
		}
	return n;
}

/* ------------------------------------------------------------------ 2-bit genome */

/* gvindex_get, gene-value-index.c:96-107: returns the base character. */
static int genome_base(const or_index_t *idx, uint32_t p)
{
	uint32_t byte = (p - idx->start_base_offset) >> 2, bit = p % 4 * 2;
	if (byte >= idx->values_bytes - 1) return 'N';
	return "AGCT"[(idx->values[byte] >> bit) & 3];
}

/* match_chro, gene-value-index.c:856-959 (base space). */
int or_match_chro(const char *read, const or_index_t *idx, uint32_t pos, int test_len, int is_negative_strand)
{
	int ret = 0, i;
	if ((uint32_t)(pos + test_len) >= idx->length + idx->start_point) return 0;
	if (pos > 0xffff0000u) return 0;
	if (is_negative_strand) {
		for (i = test_len - 1; i >= 0; i--) {
			int g = genome_base(idx, pos + test_len - 1 - i);
			switch (g) {
			case 'A': ret += read[i] == 'T'; break;
			case 'T': ret += read[i] == 'A'; break;
			case 'G': ret += read[i] == 'C'; break;
			case 'C': ret += read[i] == 'G'; break;
			}
		}
	} else {
		uint32_t byte = (pos - idx->start_base_offset) >> 2, bit = pos % 4 * 2;
		if (byte >= idx->values_bytes) return 0;
		int8_t cur = (int8_t)idx->values[byte];
		for (i = 0; i < test_len; i++) {
			int g = (cur >> bit) & 3;
			switch (read[i]) {
			case 'A': ret += g == 0; break;
			case 'G': ret += g == 1; break;
			case 'C': ret += g == 2; break;
			case 0: break;
			default: ret += g == 3;
			}
			bit += 2;
			if (bit == 8) {
				byte++;
				if (byte == idx->values_bytes) return 0;
				cur = (int8_t)idx->values[byte];
				bit = 0;
			}
		}
	}
	return ret;
}

/* ------------------------------------------------------------------ lookup + vote */

static inline uint32_t vote_row(uint32_t kv) { return (kv / SEG) % VT_ROWS; } /* _index_vote_tol */

/* gehash_go_X, sorted-hashtable.c:937-1123. */
static void go_X(const or_index_t *idx, uint32_t raw_key, int offset, int is_reversed, vote_table_t *vote,
                 int indel_tolerance, int subread_number, uint32_t low_border, uint32_t high_border, int run_round,
                 uint32_t *shift_locs, uint32_t *shift_n, or_stats_t *st)
{
	uint32_t nb = idx->buckets_number;
	int16_t key = (int16_t)(raw_key / nb);
	uint32_t b = raw_key % nb;
	const int16_t *keys = idx->keys + idx->bucket_start[b];
	const uint32_t *vals = idx->vals + idx->bucket_start[b];
	int items = (int)(idx->bucket_start[b + 1] - idx->bucket_start[b]);
	if (st) st->lookups++;
	if (!items) return;

	int imin = 0, imax = items - 1, at;
	for (;;) { /* :960-981 */
		at = (imin + imax) / 2;
		int16_t k = keys[at];
		if (st) st->probes++;
		if (k > key) imax = at - 1;
		else if (k < key) imin = at + 1;
		else break;
		if (imax < imin) return;
	}

	int sub_p1 = subread_number + 1, of_p16 = offset + 16;
	int strand_mask = is_reversed ? NEG_STRAND_MASK : 0;
	int start_at = at, scan_back = 0;
	int ii_end = SEG;
	if (indel_tolerance > 5)
		ii_end = (indel_tolerance % SEG) ? (indel_tolerance - indel_tolerance % SEG + SEG) : indel_tolerance;

	for (;;) {
		uint32_t kv = vals[at] - (uint32_t)offset;
		uint32_t home = vote_row(kv);
		int home_items = vote->items[home];
		int found = 0;
		if (st) st->hits++;

		for (int iix = 0; iix <= ii_end; iix = iix > 0 ? -iix : (-iix + SEG)) { /* 0,+5,-5,+10,-10.. :1007 */
			uint32_t row = iix ? vote_row(kv + (uint32_t)iix) : home;
			int n = vote->items[row];
			for (int i = 0; i < n; i++) {
				int dist0 = (int)(kv - vote->pos[row][i]);
				int tol = (run_round > 0 && vote->marked_shift_indel[row][i]) ? 0 : indel_tolerance;
				if (!(dist0 >= -tol && dist0 <= tol && strand_mask == vote->masks[row][i])) continue;
				int toli = vote->toli[row][i];

				if (run_round == 0 && toli > 0 && dist0 == 0 && !vote->marked_shift_indel[row][i]) { /* :1022 */
					vote->marked_shift_indel[row][i] = 1;
					shift_locs[(*shift_n)++] = vote->pos[row][i];
				}
				if (sub_p1 == vote->last_subread_cluster[row][i] && toli > 0) { /* roll-back :1027 */
					int move_dist = 0;
					if (toli >= 3) move_dist = vote->indel_recorder[row][i][toli - 3 + 2];
					int new_dist = move_dist;
					move_dist -= vote->indel_recorder[row][i][toli + 2];
					new_dist -= dist0;
					if (abs(move_dist) > abs(new_dist)) {
						toli -= 3;
						vote->toli[row][i] = (int8_t)toli;
						vote->last_subread_cluster[row][i]--;
						vote->votes[row][i]--;
					}
				}
				if (sub_p1 <= vote->last_subread_cluster[row][i]) continue; /* :1041 */

				int16_t nv = (int16_t)(vote->votes[row][i] + 1);
				vote->votes[row][i] = nv;
				if (offset + 16 > vote->coverage_end[row][i]) vote->coverage_end[row][i] = (int16_t)of_p16;
				if (dist0 == vote->current_indel_cursor[row][i]) {
					vote->indel_recorder[row][i][toli + 1] = (int16_t)sub_p1;
				} else {
					toli += 3;
					if (toli < MAX_SECTIONS * 3) {
						vote->toli[row][i] = (int8_t)toli;
						vote->indel_recorder[row][i][toli] = (int16_t)sub_p1;
						vote->indel_recorder[row][i][toli + 1] = (int16_t)sub_p1;
						vote->indel_recorder[row][i][toli + 2] = (int16_t)dist0;
						if (toli < MAX_SECTIONS * 3 - 3) vote->indel_recorder[row][i][toli + 3] = 0;
					}
					vote->current_indel_cursor[row][i] = (int8_t)dist0;
				}
				vote->last_subread_cluster[row][i] = (int16_t)sub_p1;
				if (vote->max_vote < nv) vote->max_vote = nv;
				found = 1;
				break;
			}
			if (found) break;
		}

		if (!found && kv >= low_border && kv <= high_border && home_items < VT_SPACE) { /* :1074 */
			int j = home_items;
			vote->items[home]++;
			vote->pos[home][j] = kv;
			vote->masks[home][j] = strand_mask;
			vote->votes[home][j] = 1;
			vote->toli[home][j] = 0;
			vote->marked_shift_indel[home][j] = 0;
			if (run_round > 0)
				for (uint32_t k = 0; k < *shift_n; k++)
					if (kv >= shift_locs[k] - (uint32_t)indel_tolerance && kv <= shift_locs[k] + (uint32_t)indel_tolerance) {
						vote->marked_shift_indel[home][j] = 1;
						break;
					}
			vote->indel_recorder[home][j][0] = vote->indel_recorder[home][j][1] = (int16_t)sub_p1;
			vote->indel_recorder[home][j][2] = 0;
			vote->indel_recorder[home][j][3] = 0;
			vote->current_indel_cursor[home][j] = 0;
			vote->coverage_start[home][j] = (int16_t)offset;
			vote->coverage_end[home][j] = (int16_t)of_p16;
			vote->last_subread_cluster[home][j] = (int16_t)sub_p1;
			if (vote->max_vote == 0) vote->max_vote = 1;
		}

		if (!scan_back) { /* forward to the end of the equal-key run, then backward :1109-1119 */
			at++;
			if (st && at < items) st->probes++;
			if (at == items || keys[at] != key) { scan_back = 1; at = start_at; }
		}
		if (scan_back) {
			at--;
			if (st && at >= 0) st->probes++;
			if (at < 0 || keys[at] != key) break;
		}
	}
}

/* ------------------------------------------------------------------ vote -> candidates */

/* locate_gene_position_max with head/tail_cut NULL and rl 0, gene-algorithms.c:441-510.
 * Returns 0 and (*chr,*pos) when the position is inside a chromosome body. */
static int locate_position(const or_index_t *idx, uint32_t linear, int *chr, int *pos)
{
	int n, lo = 0, hi = idx->total_offsets;
	*chr = -1;
	*pos = -1;
	for (;;) {
		if (hi <= lo + 1) { n = lo - 2 > 0 ? lo - 2 : 0; break; }
		int mid = (lo + hi) / 2;
		if (idx->read_offsets[mid] > linear) hi = mid; else lo = mid + 1;
	}
	for (; n < idx->total_offsets; n++) {
		if (idx->read_offsets[n] > linear) {
			*pos = (int)(n == 0 ? linear : linear - idx->read_offsets[n - 1]);
			if (linear > idx->read_offsets[n] + 15 - (uint32_t)idx->padding) return 1;
			if (*pos < idx->padding) return 1;
			*pos -= idx->padding;
			*chr = n;
			return 0;
		}
	}
	return -1;
}

/* test_PE_and_same_chro, core.c:4819-4845. */
static void pe_and_same_chro(const or_index_t *idx, const or_params_t *p, uint32_t pos1, uint32_t pos2, int *is_pe,
                             int *same_chr, int rlen1, int rlen2)
{
	int c1, c2, p1, p2;
	*is_pe = 0;
	*same_chr = 0;
	int r1 = locate_position(idx, pos1, &c1, &p1);
	int r2 = locate_position(idx, pos2, &c2, &p2);
	if (r1 == 0 && r2 == 0) {
		long long tlen = (long long)p1 - p2;
		if (tlen < 0) tlen = -tlen;
		tlen += p1 > p2 ? rlen1 : rlen2;
		uint32_t t = (uint32_t)tlen;
		if (c1 == c2) {
			*same_chr = 1;
			if (t >= (uint32_t)p->minimum_pair_distance && t <= (uint32_t)p->maximum_pair_distance) *is_pe = 1;
		}
	}
}

/* update_top_three, core-junction.c:908-922. */
static void update_top(const or_params_t *p, int *top, int v)
{
	if (v > top[p->top_scores - 1])
		for (int x = 0; x < p->top_scores; x++) {
			if (v > top[x]) {
				for (int y = p->top_scores - 1; y > x; y--) top[y] = top[y - 1];
				top[x] = v;
				break;
			} else if (v == top[x]) break;
		}
}

/* indel_recorder_copy, sorted-hashtable.c:1144-1164. */
static int16_t recorder_copy(int16_t *dst, const int16_t *src)
{
	int16_t all = 0;
	int i = 0;
	while (src[i] && i < REC_LEN - 2) {
		dst[i] = src[i]; i++;
		dst[i] = src[i]; i++;
		dst[i] = src[i]; all = dst[i]; i++;
	}
	dst[i] = 0;
	return all;
}

typedef struct { /* simple_mapping_t, core.h:414-424 */
	int is_vote_item, i, j;
	uint32_t pos;
	int votes;
	uint16_t read_start;
} simple_t;

typedef struct { int r1, r2; unsigned long long score; } combo_t; /* vote_combination_t core.h:426-431 */


/* ------------------------------------------------------------------ subjunc: minor half + donor site */

#define CONFIRM_WINDOW 17 /* JUNCTION_CONFIRM_WINDOW core-junction.c:3569 */

static uint32_t abs32(uint32_t x) { return x > 0x7fffffffu ? (0xffffffffu - x) + 1 : x; } /* abs32uint core-junction.c:42 */

static void genome_string2(char *buf, const or_index_t *idx, uint32_t pos) /* gvindex_get_string, len 2, forward */
{
	buf[0] = (char)genome_base(idx, pos);
	buf[1] = (char)genome_base(idx, pos + 1);
}

/* is_donor_chars_part / paired_chars_part_core(.., 0), core-junction.c:3472-3503 */
static int donor_chars(const char *c) { return (c[0] == 'G' && c[1] == 'T') || (c[0] == 'A' && c[1] == 'G') || (c[0] == 'A' && c[1] == 'C') || (c[0] == 'C' && c[1] == 'T'); }
static int ceq2(const char *c, const char *t) { return c[0] == t[0] && c[1] == t[1]; }
static int donor_pair(const char *a, const char *b)
{
	int gtag = (ceq2(a, "GT") && ceq2(b, "AG")) || (ceq2(a, "AG") && ceq2(b, "GT"));
	int ctac = (ceq2(a, "CT") && ceq2(b, "AC")) || (ceq2(a, "AC") && ceq2(b, "CT"));
	if (gtag || ctac) return ceq2(a, "CT") || ceq2(a, "GT");
	return 0;
}

/* donor_score, core-junction.c:3675-3832, for subjunc's configuration: prefer_donor_receptor_junctions,
 * check_donor_at_junctions, no fusion / long deletion, max_insertion_at_junctions 0, more_accurate_fusions. */
static int donor_score(const or_index_t *idx, const or_params_t *p, uint32_t left_pos, uint32_t right_pos, int left_indel, int right_indel,
                       int normally_arranged, int guess_start, int guess_end, const char *text, int rlen, int *split_point,
                       int *gt_ag_strand, int *donor_found, int *inserted_bases, int *small_inc, int *large_inc)
{
	int need_donor_test = p->check_donor_at_junctions ? 1 : 0;
	int mism_allowed = p->more_accurate_fusions ? 0 : 1; /* mismatch_in_between_allowd :3735 */
	*small_inc = !normally_arranged;
	*large_inc = normally_arranged;
	int most_likely = (guess_start + guess_end) / 2;
	int sel_split = -1, sel_strand = -1, best = -111111;
	char dl[3] = {0, 0, 0}, dr[3] = {0, 0, 0};
	int n_points = guess_end - guess_start;
	for (int k = 0; k < n_points; k++) {
		int sp = ((k % 2) ? -((k + 1) / 2) : ((1 + k) / 2)) + most_likely;
		int donor_ok = 0;
		if (sp > rlen - CONFIRM_WINDOW) continue;
		if (sp < CONFIRM_WINDOW) continue;
		if (!p->prefer_donor_receptor_junctions) donor_ok = 0;
		else if (normally_arranged) {
			genome_string2(dl, idx, left_pos + (uint32_t)sp + (uint32_t)left_indel);
			if (donor_chars(dl)) {
				genome_string2(dr, idx, right_pos + (uint32_t)sp + (uint32_t)right_indel - 2);
				if (donor_chars(dr)) donor_ok = donor_pair(dl, dr);
			}
		} else {
			genome_string2(dl, idx, right_pos + (uint32_t)sp + (uint32_t)left_indel);
			genome_string2(dr, idx, left_pos + (uint32_t)sp + (uint32_t)right_indel - 2);
			donor_ok = donor_chars(dl) && donor_chars(dr) && donor_pair(dl, dr);
		}
		if (!(donor_ok || !need_donor_test)) continue;
		if (normally_arranged) {
			int lm = or_match_chro(text + sp - CONFIRM_WINDOW, idx, left_pos + (uint32_t)sp - CONFIRM_WINDOW + (uint32_t)left_indel, CONFIRM_WINDOW, 0);
			if (lm > CONFIRM_WINDOW - 2) {
				int rm = or_match_chro(text + sp, idx, right_pos + (uint32_t)sp + (uint32_t)right_indel, CONFIRM_WINDOW, 0);
				if (rm >= 2 * CONFIRM_WINDOW - lm - mism_allowed) {
					int lnm = or_match_chro(text + sp, idx, left_pos + (uint32_t)sp + (uint32_t)left_indel, CONFIRM_WINDOW, 0);
					int rnm = or_match_chro(text + sp - CONFIRM_WINDOW, idx, right_pos + (uint32_t)sp + (uint32_t)right_indel - CONFIRM_WINDOW, CONFIRM_WINDOW, 0);
					if (lnm <= CONFIRM_WINDOW - 5 && rnm <= CONFIRM_WINDOW - 5) {
						int score = 100 * (donor_ok * 3000 + lm + rm - lnm - rnm);
						if (score > best) { sel_strand = (dl[0] == 'G' || dr[1] == 'G'); sel_split = sp; best = score; }
					}
				}
			}
		} else {
			int rm = or_match_chro(text + sp - CONFIRM_WINDOW, idx, right_pos + (uint32_t)right_indel + (uint32_t)sp - CONFIRM_WINDOW, CONFIRM_WINDOW, 0);
			int lm = or_match_chro(text + sp, idx, left_pos + (uint32_t)sp + (uint32_t)left_indel, CONFIRM_WINDOW, 0);
			int rnm = or_match_chro(text + sp, idx, right_pos + (uint32_t)sp + (uint32_t)right_indel, CONFIRM_WINDOW, 0);
			int lnm = or_match_chro(text + sp - CONFIRM_WINDOW, idx, left_pos + (uint32_t)left_indel + (uint32_t)sp - CONFIRM_WINDOW, CONFIRM_WINDOW, 0);
			if (lm + rm >= 2 * CONFIRM_WINDOW - mism_allowed && lnm <= CONFIRM_WINDOW - 5 && rnm <= CONFIRM_WINDOW - 5) {
				int score = 100 * (donor_ok * 3000 + lm + rm - lnm - rnm);
				if (score > best) { sel_strand = (dl[0] == 'G' || dr[1] == 'G'); sel_split = sp; best = score; }
			}
		}
	}
	if (best > 0) {
		*split_point = sel_split;
		*donor_found = best >= 290000;
		*gt_ag_strand = sel_strand;
		*inserted_bases = 0;
		return (1 + best) / 100;
	}
	return 0;
}

/* insert_big_margin_record, core-junction.c:789-813 */
static void big_margin_insert(const or_params_t *p, uint16_t *rec, unsigned char votes, short start, short end, int rlen, int is_negative)
{
	int n = p->big_margin_record_size;
	if (n < 3) return;
	uint16_t s2 = (uint16_t)(is_negative ? rlen - end : start), e2 = (uint16_t)(is_negative ? rlen - start : end);
	int x;
	for (x = 0; x < n / 3; x++)
		if (votes >= rec[x * 3]) break;
	if (x < n / 3) {
		for (int y = n - 4; y >= x * 3; y--) rec[y + 3] = rec[y];
		rec[x * 3] = votes; rec[x * 3 + 1] = s2; rec[x * 3 + 2] = e2;
	}
}

/* The do_breakpoint_detection part of copy_vote_to_alignment_res, core-junction.c:1073-1334 (no fusion,
 * no long deletion: one table holds one strand, so is_strand_jumpped is always 0). */
static void minor_half_search(const or_index_t *idx, const or_params_t *p, or_mapping_result_t *r, or_subjunc_result_t *jr,
                              const vote_table_t *v, int vi, int vj, int rlen, const char *text)
{
	for (int i = 0; i < VT_ROWS; i++)
		for (int j = 0; j < v->items[i]; j++) {
			if (i == vi && j == vj) continue;
			if (r->selected_votes < v->votes[i][j]) continue;
			long long dist = (long long)v->pos[vi][vj] - (long long)v->pos[i][j];
			/* test_junction_minor, core-junction.c:889-906 */
			if (abs((int)dist) > p->maximum_intron_length) continue; /* abs() on a long long: the reference truncates to int */
			if (v->coverage_start[vi][vj] == v->coverage_start[i][j]) continue;
			if (v->coverage_end[vi][vj] == v->coverage_end[i][j]) continue;
			if (v->coverage_start[vi][vj] > v->coverage_start[i][j]) { if (v->pos[vi][vj] < v->pos[i][j]) continue; }
			else { if (v->pos[vi][vj] > v->pos[i][j]) continue; }
			/* is_better_inner, core-junction.c:975-983 */
			int old_intron = (int)abs32(v->pos[vi][vj] - jr->minor_position), intron = (int)abs32(v->pos[vi][vj] - v->pos[i][j]);
			int mv = v->votes[i][j], mcov = v->coverage_end[i][j] - v->coverage_start[i][j];
			int jcov = jr->minor_coverage_end - jr->minor_coverage_start;
			int better = mv > jr->minor_votes || (mv == jr->minor_votes && mcov > jcov) ||
			             (mv == jr->minor_votes && mcov == jcov && intron < old_intron);
			if (!better) continue;
			int overlapped;
			if (v->coverage_start[vi][vj] > v->coverage_start[i][j]) overlapped = v->coverage_end[i][j] - v->coverage_start[vi][vj];
			else overlapped = v->coverage_end[vi][vj] - v->coverage_start[i][j];
			if (overlapped > 14) continue;
			if (abs((int)dist) < 6) continue;
			int guess_start = (v->coverage_start[vi][vj] > v->coverage_start[i][j]) ? v->coverage_end[i][j] - 8 : v->coverage_end[vi][vj] - 8;
			int guess_end = (v->coverage_start[vi][vj] < v->coverage_start[i][j]) ? v->coverage_start[i][j] + 8 : v->coverage_start[vi][vj] + 8;
			int normally_arranged = 1 != (v->coverage_start[vi][vj] > v->coverage_start[i][j]) + (v->pos[vi][vj] > v->pos[i][j]);
			int left_indel = 0, right_indel = 0, major_indels = 0, minor_indel_offset = 0;
			if (rlen > LONG_READ) {
				for (int k = 0; k < MAX_SECTIONS; k++) { if (!v->indel_recorder[vi][vj][k * 3]) break; major_indels += v->indel_recorder[vi][vj][k * 3 + 2]; }
				for (int k = 0; k < MAX_SECTIONS; k++) { if (!v->indel_recorder[i][j][k * 3]) break; minor_indel_offset += v->indel_recorder[i][j][k * 3 + 2]; }
				left_indel = v->pos[vi][vj] < v->pos[i][j] ? major_indels : minor_indel_offset;
				right_indel = 0;
			}
			int split = 0, gt_ag = 0, donor_found = 0, inserted = 0, small_inc = 0, large_inc = 0;
			uint32_t lo = v->pos[vi][vj] < v->pos[i][j] ? v->pos[vi][vj] : v->pos[i][j];
			uint32_t hi = v->pos[vi][vj] > v->pos[i][j] ? v->pos[vi][vj] : v->pos[i][j];
			int replace = donor_score(idx, p, lo, hi, left_indel, right_indel, normally_arranged, guess_start > 0 ? guess_start : 0,
			                          guess_end < rlen ? guess_end : rlen, text, rlen, &split, &gt_ag, &donor_found, &inserted, &small_inc, &large_inc);
			if (!replace) continue;
			jr->minor_position = v->pos[i][j];
			jr->minor_votes = v->votes[i][j];
			jr->minor_coverage_start = (uint16_t)v->coverage_start[i][j];
			jr->minor_coverage_end = (uint16_t)v->coverage_end[i][j];
			jr->double_indel_offset = (int8_t)((minor_indel_offset & 0xf) | ((major_indels & 0xf) << 4));
			jr->split_point = (int16_t)split;
			jr->small_side_increasing_coordinate = (int8_t)small_inc;
			jr->large_side_increasing_coordinate = (int8_t)large_inc;
			jr->indel_at_junction = (int8_t)inserted;
			r->result_flags &= ~0x3;
			if (!donor_found || gt_ag > 2) r->result_flags |= 3;
			else r->result_flags = gt_ag ? (r->result_flags | 1) : (r->result_flags & ~1);
			r->result_flags &= ~4; /* CORE_IS_STRAND_JUMPED: never in this mode */
		}
}

/* copy_vote_to_alignment_res, core-junction.c:1058-1072 (the DNA part). */
static void slot_to_result(or_mapping_result_t *r, const vote_table_t *v, int i, int j, int used_subreads)
{
	r->selected_position = v->pos[i][j];
	r->selected_votes = v->votes[i][j];
	r->indels_in_confident_coverage = (int8_t)recorder_copy(r->selected_indel_record, v->indel_recorder[i][j]);
	r->confident_coverage_end = (uint16_t)v->coverage_end[i][j];
	r->confident_coverage_start = (uint16_t)v->coverage_start[i][j];
	r->result_flags = (v->masks[i][j] & NEG_STRAND_MASK) ? CORE_NEG : 0;
	r->used_subreads_in_vote = (int16_t)used_subreads;
	r->noninformative_subreads_in_vote = 0; /* always 0 on VER2 indexes, SURVEY.md appendix A */
	r->is_fully_covered = 0;
}

/* basic_sort_run (selection sort, NOT stable), core.c:4716-4729, as merge_sort uses it for
 * <= 11 items (core.c:4736-4749); ascending by score. */
static void combo_sort(combo_t *c, int n)
{
	for (int i = 0; i < n - 1; i++) {
		int m = i;
		for (int j = i + 1; j < n; j++)
			if ((long long)c[m].score - (long long)c[j].score > 0) m = j; /* comb_sort_compare returns int diff */
		if (i != m) { combo_t t = c[i]; c[i] = c[m]; c[m] = t; }
	}
}

/* process_voting_junction_PE_topK, core-junction.c:2199-2530 (DNA mode: no fusion, no junctions). */
static void topk(const or_index_t *idx, const or_params_t *p, int is_paired, vote_table_t *vt[2], const int rlen[2],
                 const int all_subreads[2], or_mapping_result_t *res /* [mates][multi_best] */, or_subjunc_result_t *jres /* same shape or NULL */,
                 uint16_t *big_margin /* [mates][27] or NULL */, char *const text[2])
{
	int mates = 1 + is_paired, mb = p->multi_best_reads;
	int top[2][MAX_TOP];
	simple_t *simple[2];
	int n_simple[2] = {0, 0};
	simple[0] = (simple_t *)malloc(sizeof(simple_t) * p->max_vote_simples * 2);
	simple[1] = simple[0] + p->max_vote_simples;

	for (int m = 0; m < mates; m++) { /* :2214-2251 */
		memset(top[m], 0, sizeof(int) * p->top_scores);
		for (int i = 0; i < VT_ROWS; i++)
			for (int j = 0; j < vt[m]->items[i]; j++) update_top(p, top[m], vt[m]->votes[i][j]);
		for (int k = 0; k < mb; k++)
			if (res[m * mb + k].selected_votes > 0) update_top(p, top[m], res[m * mb + k].selected_votes);
	}

	for (int m = 0; m < mates; m++) { /* :2256-2316 */
		int n = 0;
		for (int lv = 0; lv < p->top_scores; lv++) {
			if (n >= p->max_vote_simples) break;
			int want = top[m][lv];
			if (want < 1 || top[m][0] - want > p->max_vote_number_cutoff) break;
			for (int i = 0; i < VT_ROWS; i++) {
				if (n >= p->max_vote_simples) break;
				for (int j = 0; j < vt[m]->items[i]; j++) {
					if (n >= p->max_vote_simples) break;
					int vv = vt[m]->votes[i][j];
					if (p->do_big_margin_filtering_for_junctions && big_margin && lv == 0 && vv >= top[m][p->top_scores - 1]) /* :2281 */
						big_margin_insert(p, big_margin + m * 27, (unsigned char)vv, vt[m]->coverage_start[i][j], vt[m]->coverage_end[i][j], rlen[m],
						                  (vt[m]->masks[i][j] & NEG_STRAND_MASK) ? 1 : 0);
					if (vv == want && vv >= p->minimum_subread_for_second_read) {
						simple[m][n].is_vote_item = 1;
						simple[m][n].i = i;
						simple[m][n].j = j;
						simple[m][n].read_start = (uint16_t)vt[m]->coverage_start[i][j];
						simple[m][n].pos = vt[m]->pos[i][j];
						simple[m][n].votes = vv;
						n++;
					}
				}
			}
			for (int k = 0; k < mb; k++) {
				if (n >= p->max_vote_simples) break;
				if (res[m * mb + k].selected_votes == want) {
					simple[m][n].is_vote_item = 0;
					simple[m][n].i = k;
					simple[m][n].j = 0;
					simple[m][n].pos = res[m * mb + k].selected_position;
					simple[m][n].votes = res[m * mb + k].selected_votes;
					simple[m][n].read_start = res[m * mb + k].confident_coverage_start;
					n++;
				}
			}
		}
		n_simple[m] = n;
	}

	combo_t *comb = (combo_t *)malloc(sizeof(combo_t) * (p->max_vote_combinations + 1));
	int n_comb = 0;
	if (is_paired) { /* :2321-2372 */
		for (int i = 0; i < n_simple[0]; i++)
			for (int j = 0; j < n_simple[1]; j++) {
				int v1 = simple[0][i].votes, v2 = simple[1][j].votes;
				int is_pe = 0, same_chr = 0;
				if ((v1 > v2 ? v1 : v2) < p->minimum_subread_for_first_read) continue;
				pe_and_same_chro(idx, p, simple[0][i].pos, simple[1][j].pos, &is_pe, &same_chr, rlen[0], rlen[1]);
				if (!is_pe && (v1 < v2 ? v1 : v2) < p->minimum_subread_for_first_read) continue;
				int weight = is_pe ? 1300 : (same_chr ? 1000 : 800); /* no exonic bitmap in scope */
				unsigned long long score = (unsigned long long)((v1 + v2) * weight);
				int t;
				for (t = 0; t < n_comb; t++)
					if (comb[t].score < score) break;
				if (t < p->max_vote_combinations) {
					int last = n_comb < p->max_vote_combinations - 1 ? n_comb : p->max_vote_combinations - 1;
					for (int mv = last; mv > t; mv--) comb[mv] = comb[mv - 1];
					comb[t].r1 = i;
					comb[t].r2 = j;
					comb[t].score = score;
					if (n_comb < p->max_vote_combinations) n_comb++;
				}
			}
	}

	or_mapping_result_t *tmp = (or_mapping_result_t *)calloc((size_t)mates * mb, sizeof(or_mapping_result_t));
	or_subjunc_result_t *jtmp = (or_subjunc_result_t *)calloc((size_t)mates * mb, sizeof(or_subjunc_result_t));
	const int junc = p->do_breakpoint_detection && jres;
	int cursor[2] = {0, 0};
#define EMIT_CANDIDATE(m, loc)                                                                                             \
	do {                                                                                                               \
		if ((loc)->is_vote_item) {                                                                                     \
			slot_to_result(&tmp[(m) * mb + cursor[m]], vt[m], (loc)->i, (loc)->j, all_subreads[m]);                    \
			if (junc) minor_half_search(idx, p, &tmp[(m) * mb + cursor[m]], &jtmp[(m) * mb + cursor[m]], vt[m], (loc)->i, (loc)->j, rlen[m], text[m]); \
		} else {                                                                                                       \
			tmp[(m) * mb + cursor[m]] = res[(m) * mb + (loc)->i];                                                      \
			if (junc) jtmp[(m) * mb + cursor[m]] = jres[(m) * mb + (loc)->i];                                          \
		}                                                                                                              \
		cursor[m]++;                                                                                                   \
	} while (0)

	if (n_comb > 0) { /* :2394-2434 */
		combo_sort(comb, n_comb);
		for (int m = 0; m < mates; m++)
			for (int c = n_comb - 1; c >= 0; c--) {
				if (cursor[m] >= mb) break;
				const simple_t *loc = &simple[m][m ? comb[c].r2 : comb[c].r1];
				int exist = 0;
				for (int j = 0; j < cursor[m]; j++)
					if (tmp[m * mb + j].selected_position == loc->pos) { exist = 1; break; }
				if (exist) continue;
				EMIT_CANDIDATE(m, loc);
			}
	} else { /* :2436-2495 */
		if (n_simple[0] == 0) res[0].noninformative_subreads_in_vote = 0;
		if (is_paired && n_simple[1] == 0) res[mb].noninformative_subreads_in_vote = 0;
		if (n_simple[0] > 0 || n_simple[1] > 0)
			for (int m = 0; m < mates; m++)
				for (int s = 0; s < n_simple[m]; s++) {
					if (cursor[m] >= mb) break;
					const simple_t *loc = &simple[m][s];
					if (loc->votes < p->minimum_subread_for_first_read) continue;
					int exist = 0;
					for (int j = 0; j < cursor[m]; j++)
						if (tmp[m * mb + j].selected_position == loc->pos) { exist = 1; break; }
					if (exist) continue;
					EMIT_CANDIDATE(m, loc);
				}
	}
#undef EMIT_CANDIDATE

	for (int m = 0; m < mates; m++) /* :2505-2527 */
		for (int k = 0; k < mb; k++) {
			if (k < cursor[m]) res[m * mb + k] = tmp[m * mb + k];
			else res[m * mb + k].selected_votes = 0;
			if (junc) {
				if (k < cursor[m]) jres[m * mb + k] = jtmp[m * mb + k];
				else jres[m * mb + k].minor_votes = 0;
			}
		}
	free(tmp);
	free(jtmp);
	free(comb);
	free(simple[0]);
}

/* ------------------------------------------------------------------ the voting driver */

typedef struct {
	vote_table_t vt[2];
	char text[2][MAX_RLEN + 32];
	uint32_t shift_locs[VT_ROWS * VT_SPACE];
	int offsets[64 * 3 + 8];
} scratch_t;

/* One fragment through do_voting's body, core.c:3106-3235. */
static void vote_fragment(const or_index_t *idx, const or_params_t *p, int is_paired, const char *const raw[2],
                          const int rlen[2], or_mapping_result_t *res, or_subjunc_result_t *jres, uint16_t *big_margin, uint16_t *verify,
                          scratch_t *sc, or_stats_t *st)
{
	int mates = 1 + is_paired, mb = p->multi_best_reads;
	int gap = idx->index_gap;
	uint32_t low_border = idx->start_base_offset;            /* core.c:3081 */
	uint32_t high_border = idx->start_base_offset + idx->length; /* core.c:3082 */
	int tol = p->max_indel_length < 16 ? p->max_indel_length : 16;  /* core.c:3059 */
	vote_table_t *vt[2] = {&sc->vt[0], &sc->vt[1]};
	int all_subreads[2] = {0, 0};

	for (int m = 0; m < mates; m++) { /* fetch_next_read_pair, core.c:1187-1199 */
		memset(sc->text[m], 0, sizeof(sc->text[m]));
		memcpy(sc->text[m], raw[m], (size_t)rlen[m]);
		if (m == 0 ? p->is_first_read_reversed : p->is_second_read_reversed) or_reverse_read(sc->text[m], rlen[m]);
		if (st) st->read_bytes += (uint64_t)rlen[m];
	}

	for (int is_reversed = 0; is_reversed < 2; is_reversed++) {
		int applied = 0;
		for (int m = 0; m < mates; m++) {
			vote_init(vt[m]); /* rlen<16: the reference leaves a stale table; we define it as empty */
			if (rlen[m] < 16) continue;
			int n_lookups = or_subread_schedule(rlen[m], gap, p->total_subreads, &applied, sc->offsets, 64 * 3 + 8);
			all_subreads[m] = applied;
			uint32_t cur_high = high_border - (uint32_t)rlen[m];
			uint32_t shift_n = 0;
			for (int round = 0; round < 2; round++) { /* core.c:3153-3186 */
				vote_init(vt[m]);
				for (int l = 0; l < n_lookups; l++) {
					int o = sc->offsets[l];
					go_X(idx, or_genekey2int(sc->text[m] + o), o, is_reversed, vt[m], tol, l / gap, low_border, cur_high, round,
					     sc->shift_locs, &shift_n, st);
				}
				if (shift_n == 0) break;
			}
		}

		char *const texts[2] = {sc->text[0], sc->text[1]};
		if (is_paired) {
			topk(idx, p, 1, vt, rlen, all_subreads, res, jres, big_margin, texts);
		} else if (vt[0]->max_vote >= p->minimum_subread_for_first_read) { /* core.c:3214 */
			topk(idx, p, 0, vt, rlen, all_subreads, res, jres, big_margin, texts);
		} else if (res[0].selected_votes < 1) { /* core.c:3217-3225 */
			res[0].noninformative_subreads_in_vote = 0;
			if (res[0].used_subreads_in_vote < applied) res[0].used_subreads_in_vote = (int16_t)applied;
		}

		if (is_reversed == 0)
			for (int m = 0; m < mates; m++) or_reverse_read(sc->text[m], rlen[m]); /* core.c:3229-3234 */
	}

	/* After the loop text[] holds the reverse complement of the fetched orientation. */
	for (int m = 0; m < mates; m++)
		for (int k = 0; k < mb; k++) {
			const or_mapping_result_t *r = &res[m * mb + k];
			if (st) st->result_bytes += sizeof(or_mapping_result_t);
			if (!verify) continue;
			if (r->selected_votes < 1) { verify[m * mb + k] = 0xffff; continue; }
			int neg = (r->result_flags & CORE_NEG) ? 1 : 0;
			if (!neg) or_reverse_read(sc->text[m], rlen[m]); /* back to the fetched orientation */
			verify[m * mb + k] = (uint16_t)or_match_chro(sc->text[m], idx, r->selected_position, rlen[m], 0);
			if (!neg) or_reverse_read(sc->text[m], rlen[m]);
			if (st) st->verify_bytes += (uint64_t)((rlen[m] + 3) / 4 + 1);
		}
}

static int vote_range(const or_index_t *idx, const or_params_t *p, const char *text, const uint64_t *off, const uint16_t *len,
                      int64_t f0, int64_t f1, int is_paired, or_mapping_result_t *out_align, or_subjunc_result_t *out_junc,
                      uint16_t *out_big_margin, uint16_t *out_verify, or_stats_t *stats)
{
	int mates = 1 + is_paired, mb = p->multi_best_reads;
	scratch_t *sc = (scratch_t *)malloc(sizeof(scratch_t));
	if (!sc) return 2;
	for (int64_t f = f0; f < f1; f++) {
		const char *raw[2] = {NULL, NULL};
		int rlen[2] = {0, 0};
		for (int m = 0; m < mates; m++) {
			raw[m] = text + off[f * mates + m];
			rlen[m] = len[f * mates + m];
			if (rlen[m] > MAX_RLEN) { free(sc); return 3; }
		}
		vote_fragment(idx, p, is_paired, raw, rlen, out_align + (size_t)f * mates * mb,
		              out_junc ? out_junc + (size_t)f * mates * mb : NULL, out_big_margin ? out_big_margin + (size_t)f * mates * 27 : NULL,
		              out_verify ? out_verify + (size_t)f * mates * mb : NULL, sc, stats);
	}
	free(sc);
	return 0;
}

static int check_params(const or_index_t *idx, const or_params_t *p)
{
	if (!idx || !p || idx->index_gap < 1 || idx->index_gap > 3) return 1;
	if (p->top_scores < 1 || p->top_scores > MAX_TOP || p->multi_best_reads < 1 || p->total_subreads < 2) return 1;
	if (p->max_vote_simples < 1 || p->max_vote_combinations < 1) return 1;
	if (p->do_breakpoint_detection && p->max_insertion_at_junctions != 0) return 9;
	return 0;
}

int or_vote_chunk(const or_index_t *idx, const or_params_t *p, const char *text, const uint64_t *off, const uint16_t *len,
                  int64_t n_frags, int is_paired, or_mapping_result_t *out_align, or_subjunc_result_t *out_junc,
                  uint16_t *out_big_margin, uint16_t *out_verify, or_stats_t *stats)
{
	int r = check_params(idx, p);
	if (r) return r;
	if (p->do_breakpoint_detection && !out_junc) return 8;
	return vote_range(idx, p, text, off, len, 0, n_frags, is_paired, out_align, out_junc, out_big_margin, out_verify, stats);
}

typedef struct {
	const or_index_t *idx; const or_params_t *p; const char *text; const uint64_t *off; const uint16_t *len;
	int64_t f0, f1; int is_paired; or_mapping_result_t *out_align; or_subjunc_result_t *out_junc; uint16_t *out_big_margin;
	uint16_t *out_verify; or_stats_t stats; int rc;
} mt_job_t;

static void *mt_entry(void *arg)
{
	mt_job_t *j = (mt_job_t *)arg;
	j->rc = vote_range(j->idx, j->p, j->text, j->off, j->len, j->f0, j->f1, j->is_paired, j->out_align, j->out_junc, j->out_big_margin,
	                   j->out_verify, &j->stats);
	return NULL;
}

int or_vote_chunk_mt(const or_index_t *idx, const or_params_t *p, const char *text, const uint64_t *off, const uint16_t *len,
                     int64_t n_frags, int is_paired, or_mapping_result_t *out_align, or_subjunc_result_t *out_junc,
                     uint16_t *out_big_margin, uint16_t *out_verify, or_stats_t *stats, int n_threads)
{
	int r = check_params(idx, p);
	if (r) return r;
	if (p->do_breakpoint_detection && !out_junc) return 8;
	if (n_threads < 1) n_threads = 1;
	if (n_threads > 256) n_threads = 256;
	pthread_t th[256];
	mt_job_t *jobs = (mt_job_t *)calloc((size_t)n_threads, sizeof(mt_job_t));
	for (int t = 0; t < n_threads; t++) {
		mt_job_t *j = &jobs[t];
		j->idx = idx; j->p = p; j->text = text; j->off = off; j->len = len; j->is_paired = is_paired;
		j->f0 = n_frags * t / n_threads; j->f1 = n_frags * (t + 1) / n_threads;
		j->out_align = out_align; j->out_junc = out_junc; j->out_big_margin = out_big_margin; j->out_verify = out_verify;
		pthread_create(&th[t], NULL, mt_entry, j);
	}
	for (int t = 0; t < n_threads; t++) {
		pthread_join(th[t], NULL);
		if (jobs[t].rc) r = jobs[t].rc;
		if (stats) {
			stats->lookups += jobs[t].stats.lookups; stats->probes += jobs[t].stats.probes; stats->hits += jobs[t].stats.hits;
			stats->read_bytes += jobs[t].stats.read_bytes; stats->verify_bytes += jobs[t].stats.verify_bytes;
			stats->result_bytes += jobs[t].stats.result_bytes;
		}
	}
	free(jobs);
	return r;
}
