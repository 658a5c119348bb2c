// fusions_hd.h -- candidate generation ("find_fusions") as data-parallel stages.
//
// Reference behaviour being reproduced: source/fusions.cpp:203-473 (find_fusions), :15-89 (predict_fusion_strands),
// :93-200 (predict_transcript_start). The reference inserts every fragment x gene1 x gene2 combination into an
// unordered_map keyed by (gene1,gene2,contig1,contig2,breakpoint1,breakpoint2,direction1,direction2) in fragment-name
// order and updates the candidate in place. Here:
//   1. every fragment emits its records (one per gene pair) into SoA columns, in name order            (emit_*_fn)
//   2. records are grouped by key with an exact hash index whose representative is the SMALLEST record index;
//      numbering the groups by that index reproduces the reference's first-insertion order            (record_key_ops)
//   3. a stable radix sort by candidate id turns groups into contiguous segments, still in name order
//   4. one thread per candidate replays its segment sequentially (walk_a_fn): label state machine, subsampling,
//      anchors, split-read lists -- the inherently order-dependent part, confined to one segment
//   5. discordant mates are bucketed by (gene1,gene2,direction1,direction2) the same way; one thread per unfiltered
//      candidate scans its bucket (walk_b_fn), then strands / splice sites / 5' gene are predicted (walk_c_fn)
#pragma once
#include "model.h"
#include "annot_hd.h"
#include "prims.h"

namespace arb {

// record meta bits
enum { RM_DIR1 = 1, RM_DIR2 = 2, RM_SWAPPED = 4, RM_SPLIT = 8, RM_EXONIC1 = 16, RM_EXONIC2 = 32 };

struct record_view {
	u32* gene1; u32* gene2; u32* contigs; i32* bp1; i32* bp2; u32* meta /* bits | filter << 8 */; u32* frag; i32* anchor1; i32* anchor2;
};

struct frag_ends { u32 a1, a2; u16 c1, c2; i32 bp1, bp2, an1, an2; u32 d1, d2; bool ex1, ex2, split, swapped; };

// the two breakpoint ends a fragment supports (fusions.cpp:221-245 split reads, :305-327 discordant mates)
ARB_HD frag_ends fragment_ends(const frag_view& f, u32 i) {
	frag_ends e;
	e.split = f.n_aln[i] == 3;
	if (e.split) {
		const u32 m = f.idx(i, MATE1), s = f.idx(i, SPLIT_READ), u = f.idx(i, SUPPLEMENTARY);
		e.a1 = s; e.a2 = u; e.c1 = f.contig[s]; e.c2 = f.contig[u];
		e.bp1 = f.fwd(s) ? f.start[s] : f.end[s]; e.bp2 = f.fwd(u) ? f.end[u] : f.start[u];
		e.d1 = f.fwd(s) ? UPSTREAM : DOWNSTREAM; e.d2 = f.fwd(u) ? DOWNSTREAM : UPSTREAM;
		e.an1 = f.fwd(m) ? f.start[m] : f.end[m]; e.an2 = f.fwd(u) ? f.start[u] : f.end[u];
	} else {
		const u32 x = f.idx(i, MATE1), y = f.idx(i, MATE2);
		e.a1 = x; e.a2 = y; e.c1 = f.contig[x]; e.c2 = f.contig[y];
		e.bp1 = f.fwd(x) ? f.end[x] : f.start[x]; e.bp2 = f.fwd(y) ? f.end[y] : f.start[y];
		e.d1 = f.fwd(x) ? DOWNSTREAM : UPSTREAM; e.d2 = f.fwd(y) ? DOWNSTREAM : UPSTREAM;
		e.an1 = f.fwd(x) ? f.start[x] : f.end[x]; e.an2 = f.fwd(y) ? f.start[y] : f.end[y];
	}
	e.ex1 = f.aflags[e.a1] & AF_EXONIC; e.ex2 = f.aflags[e.a2] & AF_EXONIC;
	e.swapped = e.c1 > e.c2 || (e.c1 == e.c2 && e.bp1 > e.bp2);
	if (e.swapped) {
		u32 t = e.a1; e.a1 = e.a2; e.a2 = t; u16 c = e.c1; e.c1 = e.c2; e.c2 = c;
		i32 b = e.bp1; e.bp1 = e.bp2; e.bp2 = b; b = e.an1; e.an1 = e.an2; e.an2 = b;
		t = e.d1; e.d1 = e.d2; e.d2 = t; bool x = e.ex1; e.ex1 = e.ex2; e.ex2 = x;
	}
	return e;
}

struct emit_count_fn {
	frag_view f; u32* cnt; const u8* owned; // owned: NULL, or 1 for the fragments whose breakpoint records this part of a multi-GPU run emits (exchange.cu)
	ARB_HD void operator()(u32 i) const {
		if (owned && !owned[i]) { cnt[i] = 0; return; }
		const bool split = f.n_aln[i] == 3;
		const u32 x = f.idx(i, split ? SPLIT_READ : MATE1), y = f.idx(i, split ? SUPPLEMENTARY : MATE2);
		cnt[i] = (u32) f.genes_cnt[x] * (u32) f.genes_cnt[y];
	}
};

struct emit_fill_fn {
	frag_view f; const u32* off; record_view r;
	ARB_HD void operator()(u32 i) const {
		if (off[i + 1] == off[i]) return; // no records (or not this part's fragment)
		const frag_ends e = fragment_ends(f, i);
		const u32 n1 = f.genes_cnt[e.a1], n2 = f.genes_cnt[e.a2];
		const u32* g1 = f.genes + f.genes_off[e.a1]; const u32* g2 = f.genes + f.genes_off[e.a2];
		const u32 meta = (e.d1 ? RM_DIR1 : 0) | (e.d2 ? RM_DIR2 : 0) | (e.swapped && e.split ? RM_SWAPPED : 0) | (e.split ? RM_SPLIT : 0) |
		                 (e.ex1 ? RM_EXONIC1 : 0) | (e.ex2 ? RM_EXONIC2 : 0) | (u32) f.filter[i] << 8;
		u32 k = off[i];
		for (u32 x = 0; x < n1; ++x) for (u32 y = 0; y < n2; ++y, ++k) {
			r.gene1[k] = g1[x]; r.gene2[k] = g2[y]; r.contigs[k] = (u32) e.c1 | (u32) e.c2 << 16; r.bp1[k] = e.bp1; r.bp2[k] = e.bp2;
			r.meta[k] = meta; r.frag[k] = i; r.anchor1[k] = e.an1; r.anchor2[k] = e.an2;
		}
	}
};

ARB_HD u64 mix64(u64 h) { h ^= h >> 31; h *= 0x9E3779B97F4A7C15ULL; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ULL; h ^= h >> 32; return h; }

struct record_key_ops { // full candidate key
	record_view r;
	ARB_HD u64 hash(u32 k) const {
		u64 h = mix64((u64) r.gene1[k] << 32 | r.gene2[k]);
		h = mix64(h ^ ((u64) (u32) r.bp1[k] << 32 | (u32) r.bp2[k]));
		return mix64(h ^ ((u64) r.contigs[k] << 2 | (r.meta[k] & 3)));
	}
	ARB_HD bool equal(u32 a, u32 b) const {
		return r.gene1[a] == r.gene1[b] && r.gene2[a] == r.gene2[b] && r.contigs[a] == r.contigs[b] && r.bp1[a] == r.bp1[b] && r.bp2[a] == r.bp2[b] && ((r.meta[a] ^ r.meta[b]) & 3) == 0;
	}
};

struct bucket_key_ops { // discordant mates by (gene1, gene2, direction1, direction2); items are positions in drec[]
	record_view r; const u32* drec;
	ARB_HD u64 hash(u32 j) const { const u32 k = drec[j]; return mix64(((u64) r.gene1[k] << 32 | r.gene2[k]) * 4 + (r.meta[k] & 3)); }
	ARB_HD bool equal(u32 a, u32 b) const { const u32 x = drec[a], y = drec[b]; return r.gene1[x] == r.gene1[y] && r.gene2[x] == r.gene2[y] && ((r.meta[x] ^ r.meta[y]) & 3) == 0; }
};
struct bucket_probe {
	record_view r; const u32* drec; u32 gene1, gene2, dirs;
	ARB_HD u64 hash() const { return mix64(((u64) gene1 << 32 | gene2) * 4 + dirs); }
	ARB_HD bool equal_item(u32 j) const { const u32 k = drec[j]; return r.gene1[k] == gene1 && r.gene2[k] == gene2 && (r.meta[k] & 3) == dirs; }
};

struct head_flag_fn { const u32* first; u32* head; ARB_HD void operator()(u32 k) const { head[k] = first[k] == k; } };
struct group_id_fn { const u32* first; const u32* head_scan; u32* id; u32* ident; ARB_HD void operator()(u32 k) const { id[k] = head_scan[first[k]]; ident[k] = k; } };
struct segment_offsets_fn { // sorted keys -> seg_off[key] = first position
	const u32* keys; u32* seg_off; u32 n; u32 n_groups;
	ARB_HD void operator()(u32 p) const { if (p == 0 || keys[p] != keys[p - 1]) seg_off[keys[p]] = p; if (p == n - 1) seg_off[n_groups] = n; }
};
struct flag_discordant_fn { record_view r; u32* flag; ARB_HD void operator()(u32 k) const { flag[k] = (r.meta[k] & RM_SPLIT) ? 0 : 1; } };
struct compact_fn { const u32* flag_scan; const u32* flag_src; u32* out; u32 n; ARB_HD void operator()(u32 k) const { if (flag_scan[k + 1] != flag_scan[k]) out[flag_scan[k]] = k; (void) flag_src; (void) n; } };

ARB_HD void expand_anchor(i32& anchor, u32 direction, i32 down_value, i32 up_value) { // fusions.cpp:276-285; 0 means "unset"
	if (direction == DOWNSTREAM) { if (down_value < anchor || anchor == 0) anchor = down_value; }
	else { if (up_value > anchor || anchor == 0) anchor = up_value; }
}

struct cand_out {
	u32* gene1; u32* gene2; u16* contig1; u16* contig2; i32* bp1; i32* bp2; u8* dir1; u8* dir2;
	u32* split_reads1; u32* split_reads2; u32* discordant_mates; u8* filter; u8* bits; u8* bits2; i32* anchor1; i32* anchor2; float* evalue;
	u32* n_list1; u32* n_list2; u32* n_listd;
};

// pass A: replay a candidate's records in name order (fusions.cpp:248-297 split reads, :330-362 discordant mates)
struct walk_a_fn {
	record_view r; const u32* perm; const u32* seg_off; cand_out c; u8* kept /* per sorted position: 0 not listed, 1 list1, 2 list2 */; u32 threshold;
	ARB_HD void operator()(u32 cand) const {
		const u32 lo = seg_off[cand], hi = seg_off[cand + 1];
		const u32 k0 = perm[lo];
		const u32 d1 = r.meta[k0] & RM_DIR1 ? 1 : 0, d2 = r.meta[k0] & RM_DIR2 ? 1 : 0;
		u32 unf[2] = {0, 0}, listed[2] = {0, 0};
		u32 bits = 0; u8 label = 0; i32 an1 = 0, an2 = 0;
		for (u32 p = lo; p < hi; ++p) {
			const u32 k = perm[p];
			const u32 m = r.meta[k];
			const u8 fl = (u8) (m >> 8);
			if (m & RM_EXONIC1) bits |= CB_EXONIC1;
			if (m & RM_EXONIC2) bits |= CB_EXONIC2;
			if (p == lo || fl == F_none || label == F_duplicates) label = fl;
			u8 keep = 0;
			if (m & RM_SPLIT) {
				const u32 s = m & RM_SWAPPED ? 1 : 0;
				if (!(unf[s] >= threshold || (fl != F_none && listed[s] >= threshold))) {
					expand_anchor(an1, d1, r.anchor1[k], r.anchor1[k]);
					expand_anchor(an2, d2, r.anchor2[k], r.anchor2[k]);
					++listed[s]; if (fl == F_none) ++unf[s];
					keep = (u8) (s + 1);
				}
			} else {
				expand_anchor(an1, d1, r.anchor1[k], r.anchor1[k]);
				expand_anchor(an2, d2, r.anchor2[k], r.anchor2[k]);
			}
			kept[p] = keep;
		}
		c.gene1[cand] = r.gene1[k0]; c.gene2[cand] = r.gene2[k0]; c.contig1[cand] = (u16) r.contigs[k0]; c.contig2[cand] = (u16) (r.contigs[k0] >> 16);
		c.bp1[cand] = r.bp1[k0]; c.bp2[cand] = r.bp2[k0]; c.dir1[cand] = (u8) d1; c.dir2[cand] = (u8) d2;
		c.split_reads1[cand] = unf[0]; c.split_reads2[cand] = unf[1]; c.discordant_mates[cand] = 0;
		c.filter[cand] = label; c.bits[cand] = (u8) bits; c.bits2[cand] = 0; c.anchor1[cand] = an1; c.anchor2[cand] = an2; c.evalue[cand] = 0;
		c.n_list1[cand] = listed[0]; c.n_list2[cand] = listed[1]; c.n_listd[cand] = 0;
	}
};

struct fill_split_lists_fn {
	record_view r; const u32* perm; const u32* seg_off; const u8* kept; const u32* list1_off; const u32* list2_off; u32* list1; u32* list2;
	ARB_HD void operator()(u32 cand) const {
		u32 w1 = list1_off[cand], w2 = list2_off[cand];
		for (u32 p = seg_off[cand]; p < seg_off[cand + 1]; ++p) {
			if (kept[p] == 1) list1[w1++] = r.frag[perm[p]];
			else if (kept[p] == 2) list2[w2++] = r.frag[perm[p]];
		}
	}
};

// pass B: discordant mates that support an unfiltered candidate (fusions.cpp:368-437). `fill` = second run that writes.
struct walk_b_fn {
	record_view r; frag_view f; annot_view an; cand_out c;
	hash_index_view bucket_table; const u32* drec; const u32* bucket_head_scan; const u32* bperm; const u32* bseg_off;
	const u32* listd_off; u32* listd; u32* need_swap; i32 max_mate_gap; u32 threshold; u32 fill;
	ARB_HD void operator()(u32 cand) const {
		if (c.filter[cand] != F_none) { if (!fill) c.n_listd[cand] = 0; return; }
		const u32 g1 = c.gene1[cand], g2 = c.gene2[cand], d1 = c.dir1[cand], d2 = c.dir2[cand];
		bucket_probe probe = {r, drec, g1, g2, (d1 ? 1u : 0u) | (d2 ? 2u : 0u)};
		const u32 slot = bucket_table.find(probe);
		if (slot == hash_index_view::EMPTY) { if (!fill) c.n_listd[cand] = 0; return; }
		const u32 b = bucket_head_scan[bucket_table.mn[slot]];
		const i32 bp1 = c.bp1[cand], bp2 = c.bp2[cand];
		const i32 overlap = (c.n_list1[cand] + c.n_list2[cand] > 0) ? 2 : max_mate_gap;
		const i32 lim1 = d1 == DOWNSTREAM ? bp1 + overlap : bp1 - overlap, lim2 = d2 == DOWNSTREAM ? bp2 + overlap : bp2 - overlap;
		const i32 g1s = an.gene_start[g1], g1e = an.gene_end[g1], g2s = an.gene_start[g2], g2e = an.gene_end[g2];
		const bool intragenic = g1 == g2 || (bp1 >= g2s - 10000 && bp1 <= g2e + 10000 && bp2 >= g1s - 10000 && bp2 <= g1e + 10000); // common.hpp:275-279
		u32 listed = 0, counted = 0;
		i32 an1 = c.anchor1[cand], an2 = c.anchor2[cand];
		u32 w = fill ? listd_off[cand] : 0;
		for (u32 p = bseg_off[b]; p < bseg_off[b + 1]; ++p) {
			const u32 k = drec[bperm[p]];
			const i32 m1 = r.bp1[k], m2 = r.bp2[k];
			if (!(((d1 == DOWNSTREAM && m1 <= lim1) || (d1 == UPSTREAM && m1 >= lim1)) && ((d2 == DOWNSTREAM && m2 <= lim2) || (d2 == UPSTREAM && m2 >= lim2)))) continue;
			if (!((!intragenic && !(m1 >= g2s && m1 <= g2e) && !(m2 >= g1s && m2 <= g1e)) || (hd_abs(bp1 - m1) <= max_mate_gap && hd_abs(bp2 - m2) <= max_mate_gap))) continue;
			const u32 frag = r.frag[k];
			const u8 fl = (u8) (r.meta[k] >> 8);
			if (fl != F_none && listed >= threshold) continue;
			if (counted >= threshold) break;
			++listed; if (fl == F_none) ++counted;
			// canonical mate order: the mate with the lower (contig, breakpoint) first (fusions.cpp:414-421)
			u32 x = f.idx(frag, MATE1), y = f.idx(frag, MATE2);
			const i32 xb = f.fwd(x) ? f.end[x] : f.start[x], yb = f.fwd(y) ? f.end[y] : f.start[y];
			const bool out_of_order = f.contig[x] > f.contig[y] || (f.contig[x] == f.contig[y] && xb > yb);
			if (out_of_order) { u32 t = x; x = y; y = t; }
			expand_anchor(an1, d1, f.start[x], f.end[x]);
			expand_anchor(an2, d2, f.start[y], f.end[y]);
			if (fill) { listd[w++] = frag; if (out_of_order) need_swap[frag] = 1; }
		}
		if (!fill) { c.n_listd[cand] = listed; }
		else { c.discordant_mates[cand] = counted; c.anchor1[cand] = an1; c.anchor2[cand] = an2; }
	}
};

struct first_fragment_fn { const u32* rec_frag; const u32* perm; const u32* seg_off; u32* out; ARB_HD void operator()(u32 c) const { out[c] = rec_frag[perm[seg_off[c]]]; } };

// exchange slots 0 and 1 of flagged fragments (all per-alignment columns)
struct swap_mates_fn {
	frag_view f; const u32* need_swap; u8* swapped;
	template <class T> ARB_HD static void sw(T* col, u32 a, u32 b) { T t = col[a]; col[a] = col[b]; col[b] = t; }
	ARB_HD void operator()(u32 i) const {
		if (!need_swap[i]) return;
		const u32 a = f.idx(i, 0), b = f.idx(i, 1);
		sw(f.contig, a, b); sw(f.start, a, b); sw(f.end, a, b); sw(f.aflags, a, b); sw(f.cigar_off, a, b); sw(f.cigar_cnt, a, b);
		sw(f.seq_off, a, b); sw(f.seq_len, a, b); sw(f.genes_off, a, b); sw(f.genes_cnt, a, b);
		swapped[i] ^= 1;
	}
};

// pass C: strand vote, splice sites, 5' gene (fusions.cpp:15-89, :443-470, :93-200)
struct walk_c_fn {
	frag_view f; annot_view an; cand_out c; const u32* list1_off; const u32* list2_off; const u32* listd_off; const u32* list1; const u32* list2; const u32* listd;
	ARB_HD void operator()(u32 cand) const {
		const u32 g1 = c.gene1[cand], g2 = c.gene2[cand], d1 = c.dir1[cand], d2 = c.dir2[cand];
		const i32 bp1 = c.bp1[cand], bp2 = c.bp2[cand];
		u32 fwd_votes = 0, rev_votes = 0;
		for (u32 p = list1_off[cand]; p < list1_off[cand + 1]; ++p) {
			const u8 af = f.aflags[f.idx(list1[p], SPLIT_READ)];
			if (!(af & AF_PRED_AMBIGUOUS)) { if (af & AF_PRED_FORWARD) ++fwd_votes; else ++rev_votes; }
		}
		for (u32 p = list2_off[cand]; p < list2_off[cand + 1]; ++p) {
			const u8 af = f.aflags[f.idx(list2[p], SUPPLEMENTARY)];
			if (!(af & AF_PRED_AMBIGUOUS)) { if (af & AF_PRED_FORWARD) ++fwd_votes; else ++rev_votes; }
		}
		for (u32 p = listd_off[cand]; p < listd_off[cand + 1]; ++p) {
			const u32 frag = listd[p];
			u32 x = f.idx(frag, MATE1), y = f.idx(frag, MATE2);
			if ((f.aflags[x] & AF_PRED_AMBIGUOUS) || f.filter[frag] == F_hairpin) continue;
			if (f.contig[x] != c.contig1[cand] || f.fwd(x) != (d1 == DOWNSTREAM)) { u32 t = x; x = y; y = t; }
			else if (f.fwd(x) == f.fwd(y)) { // same contig, same orientation: decide by proximity to the breakpoints
				const i32 xe = d1 == DOWNSTREAM ? f.end[x] : f.start[x], ye = d1 == DOWNSTREAM ? f.end[y] : f.start[y];
				const u32 dist_a = (u32) (hd_abs(bp1 - xe) + hd_abs(bp2 - ye)), dist_b = (u32) (hd_abs(bp2 - xe) + hd_abs(bp1 - ye));
				if (dist_a == dist_b) continue;
				if (dist_b < dist_a) { u32 t = x; x = y; y = t; }
			}
			if (f.aflags[x] & AF_PRED_FORWARD) ++fwd_votes; else ++rev_votes;
		}
		u32 bits = c.bits[cand] & (CB_EXONIC1 | CB_EXONIC2);
		bool strands_ambiguous = fwd_votes == rev_votes;
		bool ps1 = true, ps2 = true; // FORWARD defaults (common.hpp:251)
		if (!strands_ambiguous) { ps1 = fwd_votes > rev_votes; ps2 = (d1 == d2) ? !ps1 : ps1; }
		const bool g1fwd = an.gene_strand[g1], g2fwd = an.gene_strand[g2];
		const bool g1dummy = an.gene_flags[g1] & GF_DUMMY, g2dummy = an.gene_flags[g2] & GF_DUMMY;
		const bool ex1 = bits & CB_EXONIC1, ex2 = bits & CB_EXONIC2;
		bool sp1 = false, sp2 = false;
		const u32 n_split_listed = (list1_off[cand + 1] - list1_off[cand]) + (list2_off[cand + 1] - list2_off[cand]);
		if (n_split_listed != 0 && !strands_ambiguous) {
			sp1 = ex1 && g1fwd == ps1 && is_breakpoint_spliced(an, g1, d1, bp1);
			sp2 = ex2 && g2fwd == ps2 && is_breakpoint_spliced(an, g2, d2, bp2);
		}
		// 5' gene (fusions.cpp:93-200): 1 = gene1 starts the transcript, 0 = gene2, -1 = ambiguous
		int tstart = -1;
		const bool read_through = c.contig1[cand] == c.contig2[cand] && bp2 - bp1 < 400000 && d1 == DOWNSTREAM && d2 == UPSTREAM; // common.hpp:265-269
		const u32 split_reads = c.split_reads1[cand] + c.split_reads2[cand];
		if (sp1 || (!strands_ambiguous && !g1dummy && ps1 == g1fwd)) {
			tstart = (g1fwd == (d1 == DOWNSTREAM)) ? 1 : 0;
		} else if (sp2 || (!strands_ambiguous && !g2dummy && ps2 == g2fwd)) {
			tstart = (g2fwd == (d2 == DOWNSTREAM)) ? 0 : 1;
		} else if (!strands_ambiguous) {
			const bool out1 = ps1 == (d1 == DOWNSTREAM), out2 = ps2 == (d2 == DOWNSTREAM); // strand reads away from the junction on that side
			if (out1 && !out2) tstart = 1; else if (out2 && !out1) tstart = 0;
		} else if (!ex1 && !ex2) {
			tstart = -1;
		} else if (!ex1 && ex2) {
			if (g2fwd == (d2 == DOWNSTREAM)) tstart = 0;
			else if (split_reads == 0 && read_through) tstart = 1;
		} else if (!ex2 && ex1) {
			if (g1fwd == (d1 == DOWNSTREAM)) tstart = 1;
			else if (split_reads == 0 && read_through) tstart = 1;
		} else {
			// note the reference's operator precedence: (!dummy && fwd && DOWN) || (rev && UP)   (fusions.cpp:176-183)
			if ((!g1dummy && g1fwd && d1 == DOWNSTREAM) || (!g1fwd && d1 == UPSTREAM)) tstart = 1;
			else if ((!g2dummy && g2fwd && d2 == DOWNSTREAM) || (!g2fwd && d2 == UPSTREAM)) tstart = 0;
		}
		const bool tstart_ambiguous = tstart < 0;
		if (tstart_ambiguous) tstart = 1;
		if (!tstart_ambiguous && strands_ambiguous) { // derive strands from gene orientation (fusions.cpp:189-199)
			strands_ambiguous = false;
			if (tstart == 1) { ps1 = g1fwd; ps2 = (d1 == d2) ? !ps1 : ps1; }
			else { ps2 = g2fwd; ps1 = (d1 == d2) ? !ps2 : ps2; }
		}
		if (sp1) bits |= CB_SPLICED1; if (sp2) bits |= CB_SPLICED2;
		if (ps1) bits |= CB_PSTRAND1; if (ps2) bits |= CB_PSTRAND2;
		if (strands_ambiguous) bits |= CB_PSTRANDS_AMBIGUOUS;
		if (tstart == 1) bits |= CB_TSTART_GENE1;
		c.bits[cand] = (u8) bits; c.bits2[cand] = tstart_ambiguous ? 1 : 0;
	}
};

} // namespace arb
