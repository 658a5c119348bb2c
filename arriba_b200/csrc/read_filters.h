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

// `counters`: 64 words (word k at counters[k * stride]), one per 3-mer: occurrences in the whole read (bits 0-9), in the first aligned part (10-19) and in
// the second (20-29); reads of more than 3,000 bases would overflow the fields.
ARB_HD bool low_entropy(const frag_view& f, u32 i, float kmer_content, u32* counters, u32 stride) {
	for (u32 mate = MATE1; mate <= MATE2; ++mate) {
		const u32 a = f.idx(i, mate);
		const u32 len = f.seq_len[a];
		if (len < 3) continue;
		const u8* seq = f.sq(a);
		// aligned (not soft-clipped) parts; hard clips do not count (filter_low_entropy.cpp:40-46 tests BAM_CSOFT_CLIP only)
		const u32* c = f.cig(a); const u32 nc = f.cigar_cnt[a];
		u32 s1 = cig_op(c[0]) == C_S ? cig_len(c[0]) : 0;
		u32 e1 = len; if (cig_op(c[nc - 1]) == C_S) e1 -= cig_len(c[nc - 1]);
		u32 s2 = s1, e2 = e1;
		if (f.n_aln[i] == 3 && mate == SPLIT_READ) {
			const u32 u = f.idx(i, SUPPLEMENTARY);
			const u32* cu = f.cig(u); const u32 nu = f.cigar_cnt[u];
			s2 = cig_op(cu[0]) == C_S ? cig_len(cu[0]) : 0;
			e2 = len; if (cig_op(cu[nu - 1]) == C_S) e2 -= cig_len(cu[nu - 1]);
			if (f.fwd(u) != f.fwd(a)) { u32 ns = len - e2, ne = len - s2; s2 = ns; e2 = ne; }
		}
		const u32 max_all = kmer_threshold(len, kmer_content);
		const u32 max_1 = kmer_threshold(e1 - s1, kmer_content);
		const u32 max_2 = kmer_threshold(e2 - s2, kmer_content);
		// Counters start at 512 - threshold, so "count >= threshold" is bit 9 of the field: one AND per counted k-mer instead of three compares. A zero
		// threshold (empty aligned part) fires at the first counted k-mer, as in the reference. Reads beyond 1,500 bases take the plain compares.
		const bool biased = len <= 1500;
		const u32 start_value = biased ? (512 - max_all) | (512 - max_1) << 10 | (512 - max_2) << 20 : 0;
		const u32 reached = 1u << 9 | 1u << 19 | 1u << 29;
		for (u32 k = 0; k < 64; ++k) counters[k * stride] = start_value;
		// "pos + 1 >= s && pos < e" as one unsigned comparison: pos - (s - 1) < e - (s - 1)
		const u32 lo1 = s1 - 1, span1 = e1 + 1 > s1 ? e1 - lo1 : 0, lo2 = s2 - 1, span2 = e2 + 1 > s2 ? e2 - lo2 : 0;
		// 2-bit code of a base: T=0 G=1 C=2, anything else 3 (filter_mismappers.cpp:33-45), looked up in a 32-bit constant indexed by the nt16 code
		const u32 code2 = ~(3u << 16 | 2u << 8 | 1u << 4); // entries 8 (T), 4 (G), 2 (C) hold 0, 1, 2 -- every other entry 3
		const u32 n_words = ((len + 31) / 32) * 4;
		u32 word = nt16_word(seq, 0, n_words);
		u32 km = (code2 >> 2 * (word >> 28) & 3) << 2 | (code2 >> 2 * (word >> 24 & 15) & 3); // bases 0 and 1
		// a k-mer overlapping a counted occurrence of itself is skipped; with k=3 only the occurrences one and two positions back can overlap
		u32 km_1 = 64, km_2 = 64; // k-mers COUNTED at pos-1 and pos-2 (64 = none)
		for (u32 pos = 0; pos + 3 < len; ++pos) { // the last k-mer of the read is never examined (filter_low_entropy.cpp:77)
			const u32 b = pos + 2;
			if ((b & 7) == 0) word = nt16_word(seq, b >> 3, n_words);
			km = (km << 2 | (code2 >> 2 * (word >> (28 - 4 * (b & 7)) & 15) & 3)) & 63;
			if (km == km_1 || km == km_2) { km_2 = km_1; km_1 = 64; continue; }
			km_2 = km_1; km_1 = km;
			const u32 in1 = pos - lo1 < span1, in2 = pos - lo2 < span2;
			const u32 w = counters[km * stride] + (1u | in1 << 10 | in2 << 20);
			counters[km * stride] = w;
			if (biased ? (w & reached) != 0 : ((w & 1023) >= max_all || (w >> 10 & 1023) >= max_1 || (w >> 20) >= max_2)) return true;
		}
	}
	return false;
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

ARB_HD u8 classify_sequences(const read_filter_params& p, const frag_view& f, const annot_view& an, u32 i, u32* scratch, u32 stride) {
	u8 label = f.filter[i];
	const u32 n = f.n_aln[i];
	const u32 a0 = f.idx(i, 0), a1 = f.idx(i, 1), a2 = f.idx(i, 2);
	if (label == F_none && p.enabled(F_mismatches)) {
		const bool multi = f.fflags[i] & FF_MULTIMAPPER;
		const u32 y = n == 2 ? a1 : a2;
		const bool viral0 = an.contig_flags[f.contig[a0]] & CF_VIRAL, viraly = an.contig_flags[f.contig[y]] & CF_VIRAL;
		bool bad = !viral0 && too_many_mismatches(p, f, an, a0, f.sq(a0), f.seq_len[a0], false, multi && !viraly);
		if (!bad && !viraly) {
			if (n == 2) bad = too_many_mismatches(p, f, an, a1, f.sq(a1), f.seq_len[a1], false, multi && !viral0);
			else bad = too_many_mismatches(p, f, an, a2, f.sq(a1), f.seq_len[a1], f.fwd(a2) != f.fwd(a1), multi && !viral0);
		}
		if (bad) label = F_mismatches;
	}
	if (p.enabled(F_low_entropy)) {
		const bool examine = label == F_none || (label != F_duplicates && is_itd_shaped(f, i, p.max_itd_length));
		if (examine && low_entropy(f, i, p.max_kmer_content, scratch, stride)) label = F_low_entropy;
	}
	return label;
}

} // namespace arb
