// read_filters.h -- the read-level filter cascade as host/device per-fragment rules.
//
// One thread evaluates one fragment; the rules are evaluated in the reference's order and the first
// hit wins (arriba.cpp:327-409: every filter skips fragments whose `filter` is already set).
// Behavioural contract per rule (reference location):
//   duplicates key ............... filter_duplicates.cpp:26-47
//   uninteresting / viral contigs  filter_uninteresting_contigs.cpp:13-24, filter_viral_contigs.cpp:13-23
//   top_expressed / low_coverage viral contigs  per-contig verdicts from host/viral.cpp, per-fragment rule here
//   read_through ................. filter_proximal_read_through.cpp:15-42
//   inconsistently_clipped ....... filter_inconsistently_clipped.cpp:13-19
//   homopolymer .................. filter_homopolymer.cpp:7-14,22-54
//   small_insert_size ............ filter_small_insert_size.cpp:15-24
//   long_gap ..................... filter_long_gap.cpp:18-87
//   same_gene .................... filter_same_gene.cpp:15-45
//   hairpin ...................... filter_hairpin.cpp:8-78
//   mismatches ................... filter_mismatches.cpp:12-65,101-133 (decision table replaces :67-99, see mismatch_table.h)
//   low_entropy .................. filter_low_entropy.cpp:11-103
#pragma once
#include "model.h"
#include "annot_hd.h"

namespace arb {

struct read_filter_params {
	u32 stage_mask;              // bit per filter_id: which rules are enabled (options -f)
	u32 stage_mask_hi;           // filter ids >= 32
	u32 homopolymer_length;      // -H, default 6
	i32 min_read_through_distance; // -R, default 10000
	u32 max_overhang;            // 5 (arriba.cpp:383)
	float max_kmer_content;      // -K, default 0.6
	u32 max_itd_length;          // -l, default 100
	u32 external_duplicate_marking; // -u
	// mismatch decision table: discard[n * table_k + k], n = aligned bases compared, k = mismatches (incl. multimapper penalty)
	const u8* mismatch_table; u32 table_n; u32 table_k;
	ARB_HD bool enabled(u32 f) const { return f < 32 ? (stage_mask >> f) & 1 : (stage_mask_hi >> (f - 32)) & 1; }
};

// ------------------------------------------------------------------------------------------- duplicates
struct dup_key { u16 c1, c2; i32 p1, p2; };
ARB_HD dup_key duplicate_key(const frag_view& f, u32 i) {
	const u32 a = f.idx(i, MATE1), b = f.idx(i, f.n_aln[i] == 2 ? MATE2 : SUPPLEMENTARY);
	dup_key k;
	k.p1 = f.fwd(a) ? (i32) ((u32) f.start[a] - f.preclip(a)) : (i32) ((u32) f.end[a] + f.postclip(a));
	k.p2 = f.fwd(b) ? (i32) ((u32) f.start[b] - f.preclip(b)) : (i32) ((u32) f.end[b] + f.postclip(b));
	k.c1 = f.contig[a]; k.c2 = f.contig[b];
	if (k.p1 > k.p2) { i32 t = k.p1; k.p1 = k.p2; k.p2 = t; u16 c = k.c1; k.c1 = k.c2; k.c2 = c; } // contigs are NOT compared (filter_duplicates.cpp:42)
	return k;
}
ARB_HD u64 dup_hash(const dup_key& k) {
	u64 h = (u64) (u32) k.p1 * 0x9E3779B97F4A7C15ULL ^ ((u64) (u32) k.p2 + 0x7F4A7C15ULL) * 0xC2B2AE3D27D4EB4FULL ^ ((u64) k.c1 << 16 | k.c2) * 0x165667B19E3779F9ULL;
	h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ULL; h ^= h >> 32;
	return h;
}
ARB_HD bool dup_equal(const dup_key& a, const dup_key& b) { return a.p1 == b.p1 && a.p2 == b.p2 && a.c1 == b.c1 && a.c2 == b.c2; }

// ------------------------------------------------------------------------------------------- small helpers
ARB_HD bool is_split_read_spliced(const frag_view& f, const annot_view& an, u32 i) {
	const u32 a = f.idx(i, SPLIT_READ);
	const u32 direction = f.fwd(a) ? UPSTREAM : DOWNSTREAM;
	const i32 bp = f.fwd(a) ? f.start[a] : f.end[a];
	for (u32 k = 0; k < f.genes_cnt[a]; ++k)
		if (is_breakpoint_spliced(an, f.genes[f.genes_off[a] + k], direction, bp)) return true;
	return false;
}

ARB_HD bool window_is_homopolymer(const u8* seq, u32 from, u32 len) {
	const u32 c = nt16_at(seq, from);
	for (u32 k = 1; k < len; ++k) if (nt16_at(seq, from + k) != c) return false;
	return true;
}

// is `bp` inside (or at the right edge of) an aligned block of the alignment? contigs are not compared (filter_hairpin.cpp:8-27)
ARB_HD bool breakpoint_within_aligned_segment(const frag_view& f, i32 bp, u32 a) {
	i32 ref = f.start[a];
	const u32* c = f.cig(a);
	for (u32 k = 0; k < f.cigar_cnt[a]; ++k) {
		const u32 op = cig_op(c[k]); const i32 len = (i32) cig_len(c[k]);
		if (op == C_N || op == C_D) ref += len;
		else if (op == C_M || op == C_X || op == C_EQ) { if (bp >= ref && bp <= ref + len) return true; ref += len; }
	}
	return false;
}

// CIGAR walk against the reference: (mismatches, compared bases). `revcomp`: the read sequence is used reverse-complemented.
ARB_HD void count_mismatches(const frag_view& f, const annot_view& an, u32 a, const u8* seq, u32 seq_len, bool revcomp, u32& mismatches, u32& aligned) {
	mismatches = 0; aligned = 0;
	const u32* c = f.cig(a);
	const u32 n = f.cigar_cnt[a];
	const bool fwd = f.fwd(a);
	const u64 base = an.contig_seq_off[f.contig[a]];
	const u32 clen = an.contig_len[f.contig[a]];
	const u32* g4 = an.assembly4 ? an.assembly4 + base / 8 : 0; // contig offsets are multiples of 64 bases
	const u32 n_words = ((seq_len + 31) / 32) * 4;               // sequences are stored in 16-byte units
	i32 ref = f.start[a]; u32 rp = 0;
	for (u32 k = 0; k < n; ++k) {
		const u32 op = cig_op(c[k]), len = cig_len(c[k]);
		switch (op) {
			case C_S: case C_H:
				rp += len;
				// the clip on the breakpoint side is expected and free of charge
				if (!((k == 0 && !fwd) || (k == n - 1 && fwd))) ++mismatches;
				break;
			case C_D: ++mismatches; ref += (i32) len; break;
			case C_N: ref += (i32) len; break;
			case C_I: ++mismatches; rp += len; break;
			case C_M: case C_EQ: case C_X:
				if (g4 && ref >= 0 && (u32) ref + len <= clen && rp + len <= seq_len) { // eight bases per step on the packed reference
					for (u32 j = 0; j < len; j += 8) {
						const u32 cnt = hd_min(8u, len - j);
						const u32 valid = 0x11111111u << (4 * (8 - cnt));
						const u32 r = revcomp ? brev32(nt16_window(seq, n_words, (i32) (seq_len - 1 - (rp + j)) - 7)) : nt16_window(seq, n_words, (i32) (rp + j));
						if (revcomp) { // bit reversal complements A/C/G/T/N; ambiguity codes stay as they are in the reference (assembly.hpp:9-22): take the slow road
							const u32 s = (r & 0x55555555u) + (r >> 1 & 0x55555555u), c4 = (s & 0x33333333u) + (s >> 2 & 0x33333333u);
							if ((c4 >> 1) & ~(c4 >> 2) & valid) {
								for (u32 t = 0; t < cnt; ++t) {
									const u32 code = nt16_complement(nt16_at(seq, seq_len - 1 - (rp + j + t)));
									if (code != NT_N) { if (nt16_char(code) != an.assembly[base + (u32) ref + j + t]) ++mismatches; ++aligned; }
								}
								continue;
							}
						}
						const u32 g = packed_window(g4, (u64) (u32) ref + j);
						const u32 is_n = r & r >> 1 & r >> 2 & r >> 3, ok = valid & ~is_n;
						const u32 x = r ^ g, differs = x | x >> 1 | x >> 2 | x >> 3;
						mismatches += popc32(differs & ok); aligned += popc32(ok);
					}
					ref += (i32) len; rp += len;
					break;
				}
				for (u32 j = 0; j < len; ++j, ++ref, ++rp) {
					if (rp >= seq_len) continue; // malformed record: nothing to compare
					u32 code = revcomp ? nt16_complement(nt16_at(seq, seq_len - 1 - rp)) : nt16_at(seq, rp);
					if (code != NT_N) {
						char r = ((u32) ref < clen) ? an.assembly[base + (u32) ref] : '\0';
						if (nt16_char(code) != r) ++mismatches;
						++aligned;
					}
				}
				break;
			default: break;
		}
	}
}

ARB_HD bool too_many_mismatches(const read_filter_params& p, const frag_view& f, const annot_view& an, u32 a, const u8* seq, u32 seq_len, bool revcomp, bool multimapper_penalty) {
	u32 mm, n;
	count_mismatches(f, an, a, seq, seq_len, revcomp, mm, n);
	if (multimapper_penalty) mm += 2;
	if (n >= p.table_n) n = p.table_n - 1;   // table is sized from the longest read; defensive clamps only
	if (mm >= p.table_k) mm = p.table_k - 1;
	return p.mismatch_table[n * p.table_k + mm] != 0;
}

ARB_HD u32 kmer3(const u8* seq, u32 pos) { // T=0 G=1 C=2 else=3 per base (filter_mismappers.cpp:33-45), 3 bases
	u32 r = 0;
	for (u32 b = 0; b < 3; ++b) {
		const u32 code = nt16_at(seq, pos + b);
		r = r << 2 | (code == NT_T ? 0u : code == NT_G ? 1u : code == NT_C ? 2u : 3u);
	}
	return r;
}

ARB_HD u32 kmer_threshold(u32 len, float kmer_content) { // unsigned(len * content / 3 + 0.5), float arithmetic then double add (filter_low_entropy.cpp:67-69)
#ifdef __CUDA_ARCH__
	float q = __fdiv_rn(__fmul_rn((float) len, kmer_content), 3.0f);
#else
	volatile float prod = (float) len * kmer_content; // volatile: forbid contraction / excess precision
	float q = prod / 3.0f;
#endif
	return (u32) ((double) q + 0.5);
}

ARB_HD bool is_itd_shaped(const frag_view& f, u32 i, u32 max_itd_length) {
	if (f.n_aln[i] != 3) return false;
	const u32 s = f.idx(i, SPLIT_READ), u = f.idx(i, SUPPLEMENTARY);
	if (f.fwd(s) != f.fwd(u) || f.contig[s] != f.contig[u]) return false;
	if (f.fwd(s)) return f.start[s] < f.end[u] && f.start[s] + (i32) max_itd_length >= f.end[u];
	return f.end[s] > f.start[u] && f.end[s] <= f.start[u] + (i32) max_itd_length;
}

// ------------------------------------------------------------------------------------------- the cascade
// Evaluates rules [uninteresting_contigs .. low_entropy] for fragment i given its current label
// (F_none or F_duplicates from the duplicate pass). Returns the final label. `early` receives the label the
// fragment had after the contig rules (what estimate_fragment_length sees, read_stats.cpp:23).
// The cascade runs as two kernels: classify_head evaluates the rules that look at coordinates, CIGARs and gene sets only; the fragments that are still
// unlabelled (or ITD-shaped, see below) are queued and classify_sequences evaluates the two rules that read the bases.
ARB_HD u8 classify_head(const read_filter_params& p, const frag_view& f, const annot_view& an, u32 i, u8& early, bool& needs_sequences) {
	u8 label = f.filter[i];
	const u32 n = f.n_aln[i];
	const u32 a0 = f.idx(i, 0), a1 = f.idx(i, 1), a2 = f.idx(i, 2);

	if (label == F_none && p.enabled(F_uninteresting_contigs)) {
		for (u32 s = 0; s < n; ++s) if (!(an.contig_flags[f.contig[f.idx(i, s)]] & CF_INTERESTING)) { label = F_uninteresting_contigs; break; }
	}
	if (label == F_none && p.enabled(F_viral_contigs)) {
		bool all_viral = true;
		for (u32 s = 0; s < n; ++s) if (!(an.contig_flags[f.contig[f.idx(i, s)]] & CF_VIRAL)) { all_viral = false; break; }
		if (all_viral) label = F_viral_contigs;
	}
	// the two heuristics on viral contigs decide per contig (host/viral.cpp); a fragment goes when any of its mates lies on such a contig
	// (filter_top_expressed_viral_contigs.cpp:133-152, filter_low_coverage_viral_contigs.cpp:32-50)
	if (label == F_none && p.enabled(F_top_expressed_viral_contigs)) {
		for (u32 s = 0; s < n; ++s) { const u8 cf = an.contig_flags[f.contig[f.idx(i, s)]]; if ((cf & CF_VIRAL) && (cf & CF_VIRAL_LOW_EXPRESSION)) { label = F_top_expressed_viral_contigs; break; } }
	}
	if (label == F_none && p.enabled(F_low_coverage_viral_contigs)) {
		for (u32 s = 0; s < n; ++s) { const u8 cf = an.contig_flags[f.contig[f.idx(i, s)]]; if ((cf & CF_VIRAL) && (cf & CF_VIRAL_FOCAL_COVERAGE)) { label = F_low_coverage_viral_contigs; break; } }
	}
	early = label;

	if (label == F_none && p.enabled(F_read_through)) {
		// the mate lying upstream ("forward") and the one downstream ("reverse") in a colinear arrangement
		u32 fm, rm;
		if (n == 2) { fm = f.fwd(a0) ? a0 : a1; rm = f.fwd(a0) ? a1 : a0; }
		else { fm = f.fwd(a1) ? a2 : a1; rm = f.fwd(a1) ? a1 : a2; }
		const bool colinear = (n == 2 ? f.fwd(fm) != f.fwd(rm) : f.fwd(fm) == f.fwd(rm)) && f.contig[fm] == f.contig[rm] && f.end[fm] < f.start[rm];
		if (colinear) {
			i32 fs, fe, rs, re;
			gene_set_extent(an, f.genes + f.genes_off[fm], f.genes_cnt[fm], fs, fe);
			gene_set_extent(an, f.genes + f.genes_off[rm], f.genes_cnt[rm], rs, re);
			if (f.end[fm] >= rs - p.min_read_through_distance || f.start[rm] <= fe + p.min_read_through_distance) label = F_read_through;
		}
	}
	if (label == F_none && p.enabled(F_inconsistently_clipped) && n == 3) {
		if ((f.fwd(a0) && f.end[a0] > f.end[a1] + 3) || (!f.fwd(a0) && f.start[a0] < f.start[a1] - 3)) label = F_inconsistently_clipped;
	}
	if (label == F_none && p.enabled(F_homopolymer) && n == 3) {
		const u32 H = p.homopolymer_length, len = f.seq_len[a1];
		const u8* seq = f.sq(a1);
		bool hit = false;
		if (f.fwd(a1)) {
			const u32 pre = f.preclip(a1);
			if (pre >= H && pre <= len && H >= 2) hit = window_is_homopolymer(seq, pre - H, H);
			if (!hit && len >= pre && len - pre >= H && H >= 2) hit = window_is_homopolymer(seq, pre, H);
		} else {
			const u32 post = f.postclip(a1);
			if (post >= H && post <= len && H >= 2) hit = window_is_homopolymer(seq, len - post, H);
			if (!hit && len >= post && len - post >= H && H >= 2) hit = window_is_homopolymer(seq, len - post - H, H);
		}
		if (hit && !is_split_read_spliced(f, an, i)) label = F_homopolymer;
	}
	if (label == F_none && p.enabled(F_small_insert_size) && n == 2) {
		if (f.fwd(a0) != f.fwd(a1) && f.contig[a0] == f.contig[a1] &&
		    ((u32) hd_abs(f.start[a0] - f.start[a1]) <= p.max_overhang || (u32) hd_abs(f.end[a0] - f.end[a1]) <= p.max_overhang)) label = F_small_insert_size;
	}
	if (label == F_none && p.enabled(F_long_gap)) {
		i32 deletion = 0;
		if (n == 3 && f.contig[a1] == f.contig[a2]) {
			if (!f.fwd(a1) && !f.fwd(a2)) deletion = f.start[a2] - f.end[a1];
			else if (f.fwd(a1) && f.fwd(a2)) deletion = f.start[a1] - f.end[a2];
		}
		const bool deletion_in_range = deletion >= 700000 && deletion <= 1500000;
		for (u32 s = 0; s < n && label == F_none; ++s) {
			const u32 a = f.idx(i, s);
			const u32* c = f.cig(a); const u32 nc = f.cigar_cnt[a];
			for (u32 k = 1; k + 1 < nc; ++k) {
				if (cig_op(c[k]) != C_N || !((i32) cig_len(c[k]) >= 700000 || deletion_in_range)) continue;
				u32 left = 0, right = 0;
				for (i32 j = (i32) k - 1; j >= 0; --j) { const u32 o = cig_op(c[j]); if (cig_is_match(c[j])) left += cig_len(c[j]); else if (!(o == C_D || o == C_I || o == C_P)) break; }
				for (u32 j = k + 1; j < nc; ++j) { const u32 o = cig_op(c[j]); if (cig_is_match(c[j])) right += cig_len(c[j]); else if (!(o == C_D || o == C_I || o == C_P)) break; }
				if (left <= 15 && right <= 15) { label = F_long_gap; break; }
			}
		}
	}
	if (label == F_none && p.enabled(F_same_gene)) {
		const u32 x = n == 2 ? a0 : a1, y = n == 2 ? a1 : a2;
		if (sets_intersect(f.genes + f.genes_off[x], f.genes_cnt[x], f.genes + f.genes_off[y], f.genes_cnt[y])) {
			if (n == 2) {
				if ((f.fwd(a0) && !f.fwd(a1) && f.start[a0] <= f.end[a1]) || (!f.fwd(a0) && f.fwd(a1) && f.end[a0] >= f.start[a1])) label = F_same_gene;
			} else {
				if ((f.fwd(a1) && f.fwd(a2) && f.start[a1] >= f.end[a2]) || (!f.fwd(a1) && !f.fwd(a2) && f.end[a1] <= f.start[a2])) label = F_same_gene;
			}
		}
	}
	if (label == F_none && p.enabled(F_hairpin)) {
		const u32 x = n == 2 ? a0 : a1, y = n == 2 ? a1 : a2;
		const bool related = sets_intersect(f.genes + f.genes_off[x], f.genes_cnt[x], f.genes + f.genes_off[y], f.genes_cnt[y]) || f.contig[x] == f.contig[y];
		if (related) {
			if (n == 2) {
				const i32 b0 = f.fwd(a0) ? f.end[a0] : f.start[a0], b1 = f.fwd(a1) ? f.end[a1] : f.start[a1];
				if (breakpoint_within_aligned_segment(f, b0, a1) || breakpoint_within_aligned_segment(f, b1, a0)) label = F_hairpin;
			} else {
				const i32 bs = f.fwd(a1) ? f.start[a1] : f.end[a1], bu = f.fwd(a2) ? f.end[a2] : f.start[a2];
				if (breakpoint_within_aligned_segment(f, bs, a2) || breakpoint_within_aligned_segment(f, bu, a1) || breakpoint_within_aligned_segment(f, bu, a0)) label = F_hairpin;
			}
		}
	}
	// ITD-shaped split reads are examined for low entropy even when an earlier rule (other than duplicates) already discarded them (filter_low_entropy.cpp:17-31)
	needs_sequences = (label == F_none && (p.enabled(F_mismatches) || p.enabled(F_low_entropy))) ||
	                  (label != F_none && label != F_duplicates && p.enabled(F_low_entropy) && is_itd_shaped(f, i, p.max_itd_length));
	return label;
}

// ------------------------------------------------------------------------------------------- the sequence rules, evaluated by a GROUP of lanes per fragment
// One thread per fragment streams its own reads: every load of a warp touches 32 different cache lines and the k-mer loop is a chain of dependent shared-memory
// updates. The rules below are written for a group of `lanes` (power of two, <= 32, inside one warp) that owns one fragment: the lanes take the 8-base words of a
// read in turn (coalesced), count mismatches by popcount and 3-mers with shared-memory atomics, and combine with shuffles. With lanes == 1 (host build, tests)
// the same code is the sequential rule.
struct lane_group {
	u32 lane, lanes;
	u32 mask; // device: the group's lanes inside the warp
	ARB_HD u32 sum(u32 v) const {
#ifdef __CUDA_ARCH__
		for (u32 d = lanes >> 1; d; d >>= 1) v += __shfl_xor_sync(mask, v, d);
#endif
		return v;
	}
	ARB_HD bool any(bool p) const {
#ifdef __CUDA_ARCH__
		return (__ballot_sync(mask, p) & mask) != 0;
#else
		return p;
#endif
	}
	ARB_HD void sync() const {
#ifdef __CUDA_ARCH__
		__syncwarp(mask);
#endif
	}
};
ARB_HD void group_add(u32* p, u32 v) { // counter shared by the lanes of a group
#ifdef __CUDA_ARCH__
	atomicAdd(p, v);
#else
	*p += v;
#endif
}

// 2-bit codes (T=0 G=1 C=2, anything else 3; filter_mismappers.cpp:33-45) of the eight nt16 bases of a word, first base in bits 15..14
ARB_HD u32 nt16_dense2(u32 w) {
	const u32 a = w >> 3, b = w >> 2, c = w >> 1, m = 0x11111111u;
	const u32 nd = ~w & m;                                  // bit 0 of the nibble clear
	const u32 is_t = a & ~b & ~c & nd, is_g = ~a & b & ~c & nd, is_c = ~a & ~b & c & nd; // nibble == 8, 4, 2
	const u32 lo = ~(is_t | is_c) & m, hi = ~(is_t | is_g) & m;
	u32 x = lo | hi << 1;                                    // one code per nibble
	x = (x | x >> 2) & 0x0f0f0f0fu; x = (x | x >> 4) & 0x00ff00ffu; x = (x | x >> 8) & 0xffffu;
	return x;
}

// CIGAR walk against the reference by a group: the lanes split every aligned block into 8-base chunks. Returns the same (mismatches, compared bases) as count_mismatches.
ARB_HD void count_mismatches_group(const lane_group& g, const frag_view& f, const annot_view& an, u32 a, const u32* c /* CIGAR of a */, const u8* seq, u32 seq_len, bool revcomp, u32& mismatches, u32& aligned) {
	u32 mm_all = 0, mm = 0, al = 0; // counted once / by this lane
	const u32 n = f.cigar_cnt[a];
	const bool fwd = f.fwd(a);
	const u64 base = an.contig_seq_off[f.contig[a]];
	const u32 clen = an.contig_len[f.contig[a]];
	const u32* g4 = an.assembly4 ? an.assembly4 + base / 8 : 0;
	const u32 n_words = ((seq_len + 31) / 32) * 4;
	i32 ref = f.start[a]; u32 rp = 0;
	for (u32 k = 0; k < n; ++k) {
		const u32 op = cig_op(c[k]), len = cig_len(c[k]);
		switch (op) {
			case C_S: case C_H:
				rp += len;
				if (!((k == 0 && !fwd) || (k == n - 1 && fwd))) ++mm_all;
				break;
			case C_D: ++mm_all; ref += (i32) len; break;
			case C_N: ref += (i32) len; break;
			case C_I: ++mm_all; rp += len; break;
			case C_M: case C_EQ: case C_X:
				if (g4 && ref >= 0 && (u32) ref + len <= clen && rp + len <= seq_len) {
					for (u32 j = 8 * g.lane; j < len; j += 8 * g.lanes) {
						const u32 cnt = hd_min(8u, len - j);
						const u32 valid = 0x11111111u << (4 * (8 - cnt));
						const u32 r = revcomp ? brev32(nt16_window(seq, n_words, (i32) (seq_len - 1 - (rp + j)) - 7)) : nt16_window(seq, n_words, (i32) (rp + j));
						if (revcomp) { // ambiguity codes are not complemented by the reference (assembly.hpp:9-22): base by base
							const u32 s = (r & 0x55555555u) + (r >> 1 & 0x55555555u), c4 = (s & 0x33333333u) + (s >> 2 & 0x33333333u);
							if ((c4 >> 1) & ~(c4 >> 2) & valid) {
								for (u32 t = 0; t < cnt; ++t) {
									const u32 code = nt16_complement(nt16_at(seq, seq_len - 1 - (rp + j + t)));
									if (code != NT_N) { if (nt16_char(code) != an.assembly[base + (u32) ref + j + t]) ++mm; ++al; }
								}
								continue;
							}
						}
						const u32 gw = packed_window(g4, (u64) (u32) ref + j);
						const u32 is_n = r & r >> 1 & r >> 2 & r >> 3, ok = valid & ~is_n;
						const u32 x = r ^ gw, differs = x | x >> 1 | x >> 2 | x >> 3;
						mm += popc32(differs & ok); al += popc32(ok);
					}
				} else {
					for (u32 j = g.lane; j < len; j += g.lanes) {
						if (rp + j >= seq_len) continue;
						const u32 code = revcomp ? nt16_complement(nt16_at(seq, seq_len - 1 - (rp + j))) : nt16_at(seq, rp + j);
						if (code != NT_N) {
							const i32 q = ref + (i32) j;
							const char r = ((u32) q < clen) ? an.assembly[base + (u32) q] : '\0';
							if (nt16_char(code) != r) ++mm;
							++al;
						}
					}
				}
				ref += (i32) len; rp += len;
				break;
			default: break;
		}
	}
	mismatches = mm_all + g.sum(mm); aligned = g.sum(al);
}

ARB_HD bool too_many_mismatches_group(const lane_group& g, const read_filter_params& p, const frag_view& f, const annot_view& an, u32 a, const u32* cig, const u8* seq, u32 seq_len, bool revcomp, bool multimapper_penalty) {
	u32 mm, n;
	count_mismatches_group(g, f, an, a, cig, seq, seq_len, revcomp, mm, n);
	if (multimapper_penalty) mm += 2;
	if (n >= p.table_n) n = p.table_n - 1;
	if (mm >= p.table_k) mm = p.table_k - 1;
	return p.mismatch_table[n * p.table_k + mm] != 0;
}

// scopes of the 3-mer counts of one mate (filter_low_entropy.cpp:40-69): the whole read and two aligned windows
struct entropy_scopes { u32 len, s1, e1, s2, e2, max_all, max_1, max_2; };
ARB_HD entropy_scopes entropy_scopes_of(const frag_view& f, u32 i, u32 mate, const u32* c /* CIGAR of the mate */, const u32* cu /* of the supplementary */, float kmer_content) {
	entropy_scopes sc;
	const u32 a = f.idx(i, mate);
	const u32 len = f.seq_len[a]; const u32 nc = f.cigar_cnt[a];
	sc.len = len;
	sc.s1 = cig_op(c[0]) == C_S ? cig_len(c[0]) : 0;
	sc.e1 = len; if (cig_op(c[nc - 1]) == C_S) sc.e1 -= cig_len(c[nc - 1]);
	sc.s2 = sc.s1; sc.e2 = sc.e1;
	if (f.n_aln[i] == 3 && mate == SPLIT_READ) {
		const u32 u = f.idx(i, SUPPLEMENTARY); const u32 nu = f.cigar_cnt[u];
		sc.s2 = cig_op(cu[0]) == C_S ? cig_len(cu[0]) : 0;
		sc.e2 = len; if (cig_op(cu[nu - 1]) == C_S) sc.e2 -= cig_len(cu[nu - 1]);
		if (f.fwd(u) != f.fwd(a)) { const u32 ns = len - sc.e2, ne = len - sc.s2; sc.s2 = ns; sc.e2 = ne; }
	}
	sc.max_all = kmer_threshold(len, kmer_content); sc.max_1 = kmer_threshold(sc.e1 - sc.s1, kmer_content); sc.max_2 = kmer_threshold(sc.e2 - sc.s2, kmer_content);
	return sc;
}

// the reference's sequential count of one mate: identical 3-mers that overlap a counted occurrence are skipped (filter_low_entropy.cpp:74-103)
ARB_HD bool low_entropy_mate_exact(const entropy_scopes& sc, const u8* seq, u32* counters, u32 stride) {
	const u32 len = sc.len;
	const bool biased = len <= 1500;
	const u32 start_value = biased ? (512 - sc.max_all) | (512 - sc.max_1) << 10 | (512 - sc.max_2) << 20 : 0;
	const u32 reached = 1u << 9 | 1u << 19 | 1u << 29;
	for (u32 k = 0; k < 64; ++k) counters[k * stride] = start_value;
	const u32 lo1 = sc.s1 - 1, span1 = sc.e1 + 1 > sc.s1 ? sc.e1 - lo1 : 0, lo2 = sc.s2 - 1, span2 = sc.e2 + 1 > sc.s2 ? sc.e2 - lo2 : 0;
	const u32 code2 = ~(3u << 16 | 2u << 8 | 1u << 4);
	const u32 n_words = ((len + 31) / 32) * 4;
	u32 word = nt16_word(seq, 0, n_words);
	u32 km = (code2 >> 2 * (word >> 28) & 3) << 2 | (code2 >> 2 * (word >> 24 & 15) & 3);
	u32 km_1 = 64, km_2 = 64;
	for (u32 pos = 0; pos + 3 < len; ++pos) {
		const u32 b = pos + 2;
		if ((b & 7) == 0) word = nt16_word(seq, b >> 3, n_words);
		km = (km << 2 | (code2 >> 2 * (word >> (28 - 4 * (b & 7)) & 15) & 3)) & 63;
		if (km == km_1 || km == km_2) { km_2 = km_1; km_1 = 64; continue; }
		km_2 = km_1; km_1 = km;
		const u32 in1 = pos - lo1 < span1, in2 = pos - lo2 < span2;
		const u32 w = counters[km * stride] + (1u | in1 << 10 | in2 << 20);
		counters[km * stride] = w;
		if (biased ? (w & reached) != 0 : ((w & 1023) >= sc.max_all || (w >> 10 & 1023) >= sc.max_1 || (w >> 20) >= sc.max_2)) return true;
	}
	return false;
}

// Group version, exact. Whether the k-mer at position p is counted depends only on the two positions before it:
//   counted[p] = !((km[p] == km[p-1] && counted[p-1]) || (km[p] == km[p-2] && counted[p-2]))
// and km[p] == km[p-1] needs four equal bases in a row, km[p] == km[p-2] a period-2 stretch of five: in ordinary reads 30 of 32 positions are counted
// unconditionally. The lanes take the 8-position words of the read in turn. Pass A stores per word which positions equal a predecessor; in pass B a lane
// walks back to the nearest word without such positions (after it everything is counted), replays the few words in between to learn the state at its own
// word, and adds its counted k-mers to the group's 64 counters with shared-memory atomics. Counters start at 512 - threshold per scope (three 10-bit fields),
// so "some scope reached its threshold" is one OR over the counters at the end -- counts only grow, the reference's early exit sees the same.
// `counters`: 64 words of the group, `dense`: two words per 8 bases of the longest read (+2): 2-bit codes, then the predecessor masks.
ARB_HD void entropy_replay_word(u32 m, u32& c1, u32& c2, u32& counted) { // m: bit 14-2j = equals predecessor at distance 1, bit 30-2j = at distance 2 (position j of the word)
	counted = 0;
	#pragma unroll
	for (u32 j = 0; j < 8; ++j) {
		const u32 e1 = m >> (14 - 2 * j) & 1u, e2 = m >> (30 - 2 * j) & 1u;
		const u32 c = ((e1 & c1) | (e2 & c2)) ^ 1u;
		counted |= c << j; c2 = c1; c1 = c;
	}
}
ARB_HD bool low_entropy_mate_group(const lane_group& g, const entropy_scopes& sc, const u8* seq, u32* counters, u32* dense) {
	const u32 len = sc.len;
	if (len < 4) return false; // no position is examined (pos + 3 < len)
	if (len > 1000) { // a 10-bit field (512 - threshold + count) could wrap: the plain sequential count
		bool hit = false;
		if (g.lane == 0) hit = low_entropy_mate_exact(sc, seq, counters, 1);
		hit = g.any(hit); g.sync();
		return hit;
	}
	const u32 start_value = (512 - sc.max_all) | (512 - sc.max_1) << 10 | (512 - sc.max_2) << 20, reached = 1u << 9 | 1u << 19 | 1u << 29;
	const u32 n_words = ((len + 31) / 32) * 4, used_words = (len + 7) / 8;
	const i32 last = (i32) len - 3; // positions 0 .. last-1 are examined (the last k-mer never is, filter_low_entropy.cpp:77)
	const u32 W = ((u32) last + 7) / 8;
	u32* const masks = dense + used_words + 2;
	for (u32 k = g.lane; k < 64; k += g.lanes) counters[k] = start_value;
	for (u32 w = g.lane; w <= used_words; w += g.lanes) dense[w] = nt16_dense2(nt16_word(seq, w, n_words));
	g.sync();
	for (u32 w = g.lane; w < W; w += g.lanes) { // pass A: which positions repeat the k-mer one or two positions before
		const u64 d = (u64) (w ? dense[w - 1] : 0u) << 32 | (u64) dense[w] << 16 | dense[w + 1]; // bases 8w-8 .. 8w+15, base t of the window in bits 47-2t, 46-2t
		const u64 x1 = d ^ d >> 2, x2 = d ^ d >> 4;                                                // field t: base t against base t-1 / t-2
		const u64 z1 = ~(x1 | x1 >> 1) & 0x555555555555ull, z2 = ~(x2 | x2 >> 1) & 0x555555555555ull; // bit 46-2t: equal
		u32 e1 = (u32) ((z1 & z1 << 2 & z1 << 4) >> 16) & 0x5555u, e2 = (u32) ((z2 & z2 << 2 & z2 << 4) >> 16) & 0x5555u; // position j of the word (t = 8+j) at bit 14-2j
		if (w == 0) { e1 &= 0x1555u; e2 &= 0x0555u; } // positions 0 (and 1) have no predecessor at that distance
		masks[w] = e1 | e2 << 16;
	}
	g.sync();
	// positions of window x: a_x <= pos < b_x (pos + 1 >= s && pos < e)
	const i32 a1 = (i32) hd_max(sc.s1, 1u) - 1, b1 = sc.e1 + 1 > sc.s1 ? (i32) sc.e1 : 0, a2 = (i32) hd_max(sc.s2, 1u) - 1, b2 = sc.e2 + 1 > sc.s2 ? (i32) sc.e2 : 0;
	for (u32 w = g.lane; w < W; w += g.lanes) { // pass B
		u32 ws = w;
		while (ws > 0 && masks[ws - 1] != 0) --ws;
		u32 c1 = ws ? 1u : 0u, c2 = c1, counted;
		for (u32 u = ws; u < w; ++u) entropy_replay_word(masks[u], c1, c2, counted);
		const u32 m = masks[w];
		if (m) entropy_replay_word(m, c1, c2, counted); else counted = 0xffu;
		const i32 p0 = (i32) (8 * w);
		const u32 d = dense[w] << 16 | dense[w + 1]; // bases p0 .. p0+15
		counted &= (1u << (u32) hd_min(8, last - p0)) - 1u; // examined positions of this word
		#define ARB_RANGE_MASK(a_, b_) ((((1u << (u32) hd_min(hd_max((b_) - p0, 0), 8)) - 1u) & ~((1u << (u32) hd_min(hd_max((a_) - p0, 0), 8)) - 1u)))
		const u32 in = ARB_RANGE_MASK(a1, b1) | ARB_RANGE_MASK(a2, b2) << 10; // bit j: position p0+j in window 1, bit 10+j: in window 2
		#undef ARB_RANGE_MASK
		#pragma unroll
		for (u32 j = 0; j < 8; ++j)
			if (counted >> j & 1u) group_add(&counters[d >> (26 - 2 * j) & 63u], 1u + ((in >> j & 0x401u) << 10));
	}
	g.sync();
	u32 acc = 0;
	for (u32 k = g.lane; k < 64; k += g.lanes) acc |= counters[k];
	const bool low = g.any((acc & reached) != 0);
	g.sync();
	return low;
}

// pointers of the fragment's data as the kernel staged them (shared memory) or where they lie (global memory)
struct fragment_inputs { const u8* seq0; const u8* seq1; const u32* cig0; const u32* cig1; const u32* cig2; };
ARB_HD fragment_inputs fragment_inputs_of(const frag_view& f, u32 i) {
	fragment_inputs in;
	const u32 a0 = f.idx(i, 0), a1 = f.idx(i, 1), a2 = f.idx(i, 2);
	in.seq0 = f.sq(a0); in.seq1 = f.sq(a1); in.cig0 = f.cig(a0); in.cig1 = f.cig(a1); in.cig2 = f.cig(a2);
	return in;
}

ARB_HD u8 classify_sequences_group(const lane_group& g, const read_filter_params& p, const frag_view& f, const annot_view& an, u32 i, const fragment_inputs& in, u32* counters, u32* dense) {
	u8 label = f.filter[i];
	const u32 n = f.n_aln[i];
	const u32 a0 = f.idx(i, 0), a1 = f.idx(i, 1), a2 = f.idx(i, 2);
	if (label == F_none && p.enabled(F_mismatches)) {
		const bool multi = f.fflags[i] & FF_MULTIMAPPER;
		const u32 y = n == 2 ? a1 : a2;
		const bool viral0 = an.contig_flags[f.contig[a0]] & CF_VIRAL, viraly = an.contig_flags[f.contig[y]] & CF_VIRAL;
		bool bad = !viral0 && too_many_mismatches_group(g, p, f, an, a0, in.cig0, in.seq0, f.seq_len[a0], false, multi && !viraly);
		if (!bad && !viraly) {
			if (n == 2) bad = too_many_mismatches_group(g, p, f, an, a1, in.cig1, in.seq1, f.seq_len[a1], false, multi && !viral0);
			else bad = too_many_mismatches_group(g, p, f, an, a2, in.cig2, in.seq1, f.seq_len[a1], f.fwd(a2) != f.fwd(a1), multi && !viral0);
		}
		if (bad) label = F_mismatches;
	}
	if (p.enabled(F_low_entropy)) {
		const bool examine = label == F_none || (label != F_duplicates && is_itd_shaped(f, i, p.max_itd_length));
		if (examine) {
			bool low = low_entropy_mate_group(g, entropy_scopes_of(f, i, MATE1, in.cig0, in.cig2, p.max_kmer_content), in.seq0, counters, dense);
			if (!low) low = low_entropy_mate_group(g, entropy_scopes_of(f, i, MATE2, in.cig1, in.cig2, p.max_kmer_content), in.seq1, counters, dense);
			if (low) label = F_low_entropy;
		}
	}
	return label;
}

// one thread per fragment (host build, and the reference point of the tests)
ARB_HD u8 classify_sequences(const read_filter_params& p, const frag_view& f, const annot_view& an, u32 i, u32* scratch, u32 stride) {
	(void) stride; // scratch: 64 counters + dense codes, contiguous (the one-lane group has no bank conflicts to avoid)
	lane_group g; g.lane = 0; g.lanes = 1; g.mask = 0;
#ifdef __CUDA_ARCH__
	g.mask = 1u << (threadIdx.x & 31u);
#endif
	return classify_sequences_group(g, p, f, an, i, fragment_inputs_of(f, i), scratch, scratch + 64);
}

} // namespace arb
