// events_hd.h -- candidate-level ("event-level") device stages: merging of adjacent breakpoints and the e-value model.
//
// Reference behaviour: merge_adjacent_fusions (merge_adjacent_fusions.cpp:19-108), estimate_expected_fusions
// (filter_relative_support.cpp:17-207), filter_relative_support (:209-223).
//
// merge_adjacent is sequential in coordinate order and updates the very counters it compares, but two candidates can only
// interact when their first breakpoints lie within max_distance on the same contig. After a radix sort by
// (contig1, breakpoint1, contig2, breakpoint2) the table therefore splits into independent clusters (gap > max_distance
// or contig change); one thread replays one cluster sequentially, clusters run in parallel.
#pragma once
#include "model.h"
#include "annot_hd.h"
#include "prims.h"

namespace arb {

struct cand_state { // mutable event-level columns, device copies
	u32 n;
	const u32* gene1; const u32* gene2; const u16* contig1; const u16* contig2; const i32* bp1; const i32* bp2; const u8* dir1; const u8* dir2; const u8* bits;
	u32* split_reads1; u32* split_reads2; u32* discordant_mates; u8* filter; float* evalue;
	u32* n_list1; u32* n_list2; // current sizes of the split-read lists (grow when ITD lists are merged)
};

ARB_HD bool cand_is_itd(const cand_state& c, u32 k, u32 max_itd_length) { // common.hpp:270-274
	return c.gene1[k] == c.gene2[k] && ((u32) c.bp2[k] - (u32) c.bp1[k]) < max_itd_length && c.dir1[k] == UPSTREAM && c.dir2[k] == DOWNSTREAM;
}
ARB_HD u32 cand_support(const cand_state& c, u32 k) { return c.split_reads1[k] + c.split_reads2[k] + c.discordant_mates[k]; }

struct merge_eligible_fn { // candidates that take part: unfiltered, or ITD-shaped (merge_adjacent_fusions.cpp:23-26)
	cand_state c; u32 max_itd_length; u32* flag;
	ARB_HD void operator()(u32 k) const { flag[k] = (c.filter[k] == F_none || cand_is_itd(c, k, max_itd_length)) ? 1 : 0; }
};
struct merge_gather_fn { const u32* flag_scan; u32* ids; ARB_HD void operator()(u32 k) const { if (flag_scan[k + 1] != flag_scan[k]) ids[flag_scan[k]] = k; } };
// sort keys, least significant first: bp2, contig2, bp1, contig1 (stable LSD passes on the same permutation)
struct merge_key_fn {
	cand_state c; const u32* ids; u32* key; int which;
	ARB_HD void operator()(u32 j) const {
		const u32 k = ids[j];
		key[j] = which == 0 ? (u32) c.bp2[k] : which == 1 ? (u32) c.contig2[k] : which == 2 ? (u32) c.bp1[k] : (u32) c.contig1[k];
	}
};
struct merge_cluster_head_fn { // position j starts a cluster if it cannot interact with position j-1
	cand_state c; const u32* ids; u32* head; i32 max_distance;
	ARB_HD void operator()(u32 j) const {
		if (j == 0) { head[j] = 1; return; }
		const u32 a = ids[j - 1], b = ids[j];
		head[j] = (c.contig1[a] != c.contig1[b] || c.bp1[b] - c.bp1[a] > max_distance) ? 1 : 0;
	}
};
struct merge_cluster_start_fn { const u32* head_scan; u32* start; u32 n; u32 n_clusters; ARB_HD void operator()(u32 j) const { if (head_scan[j + 1] != head_scan[j]) start[head_scan[j]] = j; if (j == n - 1) start[n_clusters] = n; } };

struct merge_cluster_fn {
	cand_state c; const u32* ids; const u32* cluster_start; i32 max_distance; u32 max_itd_length;
	u32* itd_log; u32* itd_log_count; u32 itd_log_capacity; // (winner, loser, sorted position of winner) triples
	ARB_HD bool partner(u32 f, u32 o, bool itd, bool o_is_before) const {
		if (!(c.gene1[o] == c.gene1[f] && c.gene2[o] == c.gene2[f] && c.dir1[o] == c.dir1[f] && c.dir2[o] == c.dir2[f] && c.contig2[o] == c.contig2[f])) return false;
		const i32 sign = (c.dir1[f] == c.dir2[f]) ? +1 : -1;
		const i32 shift = o_is_before ? (c.bp1[f] - c.bp1[o]) * sign : (c.bp1[o] - c.bp1[f]) * -sign; // breakpoints must be shifted in the same sense
		if (!(c.bp2[o] == c.bp2[f] + shift || (itd && hd_abs(c.bp2[f] - c.bp2[o]) <= max_distance))) return false;
		return c.split_reads1[o] + c.split_reads2[o] > 0 || (itd && c.n_list1[o] + c.n_list2[o] > 0);
	}
	ARB_HD void operator()(u32 cl) const {
		const u32 lo = cluster_start[cl], hi = cluster_start[cl + 1];
		for (u32 j = lo; j < hi; ++j) {
			const u32 f = ids[j];
			const bool itd = cand_is_itd(c, f, max_itd_length);
			if ((!itd && c.split_reads1[f] + c.split_reads2[f] == 0) || (itd && c.n_list1[f] + c.n_list2[f] == 0)) continue;
			// neighbours: upstream (descending), then downstream (ascending), as the reference collects them; decide + sum in that order
			u32 sum1 = 0, sum2 = 0; bool most = true;
			for (int pass = 0; pass < 2 && most; ++pass) {
				if (pass == 0) {
					for (u32 q = j; q-- > lo;) {
						const u32 o = ids[q];
						if (c.bp1[o] < c.bp1[f] - max_distance) break;
						if (!partner(f, o, itd, true)) continue;
						if (cand_support(c, f) < cand_support(c, o) || (cand_support(c, f) == cand_support(c, o) && c.n_list1[f] + c.n_list2[f] < c.n_list1[o] + c.n_list2[o])) { most = false; break; }
						sum1 += c.split_reads1[o]; sum2 += c.split_reads2[o];
					}
				} else {
					for (u32 q = j + 1; q < hi; ++q) {
						const u32 o = ids[q];
						if (c.bp1[o] > c.bp1[f] + max_distance) break;
						if (!partner(f, o, itd, false)) continue;
						if (cand_support(c, f) < cand_support(c, o) || (cand_support(c, f) == cand_support(c, o) && c.n_list1[f] + c.n_list2[f] < c.n_list1[o] + c.n_list2[o])) { most = false; break; }
						sum1 += c.split_reads1[o]; sum2 += c.split_reads2[o];
					}
				}
			}
			if (!most) continue;
			c.split_reads1[f] += sum1; c.split_reads2[f] += sum2;
			for (int pass = 0; pass < 2; ++pass) {
				if (pass == 0) {
					for (u32 q = j; q-- > lo;) { const u32 o = ids[q]; if (c.bp1[o] < c.bp1[f] - max_distance) break; if (partner(f, o, itd, true)) absorb(f, o, itd, j); }
				} else {
					for (u32 q = j + 1; q < hi; ++q) { const u32 o = ids[q]; if (c.bp1[o] > c.bp1[f] + max_distance) break; if (partner(f, o, itd, false)) absorb(f, o, itd, j); }
				}
			}
		}
	}
	ARB_HD void absorb(u32 f, u32 o, bool itd, u32 pos) const {
		c.filter[o] = F_merge_adjacent;
		if (itd) { // discarded reads matter for ITDs: the loser's lists are appended to the winner's (done by the caller from this log)
			const u32 at = atomic_add_u32(itd_log_count, 1);
			if (at < itd_log_capacity) { itd_log[3 * at] = f; itd_log[3 * at + 1] = o; itd_log[3 * at + 2] = pos; }
			c.n_list1[f] += c.n_list1[o]; c.n_list2[f] += c.n_list2[o];
		}
	}
};

// ------------------------------------------------------------------------------------------- e-value
struct evalue_inputs { // host-computed, order-dependent global statistics (filter_relative_support.cpp:19-127) and pow() tables
	const i32* partner_count;      // per gene: fusion_partner_count (0 if absent)
	u32 spliced_breakpoints, exonic_breakpoints, intronic_breakpoints, exonic_intronic_breakpoints;
	u32 intragenic_duplications, intragenic_inversions, spliced_same_gene, spliced_different_genes;
	float read_through_fraction;
	u64 mapped_reads;
	const double* pow_reads;       // pow(0.02, n - 2) as the reference evaluates it, n = supporting reads (index clamped to table size)
	const double* pow_intragenic;  // pow(n - 0.42, -2.11) * pow(10, -1.11)
	const double* pow_intergenic;  // pow(n - 0.73, -2.28) * pow(10, -1.75)
	u32 n_read_table;
	const double* pow_spliced1000; // pow(max(400, d) / 1000.0, -2), d < 1000
	const double* pow_spliced400;  // pow(max(1, d) / 400.0, -4.58), d < 400
	const double* pow_read_through;// pow(max(1, d) / 400000.0, -0.63), d < 400000
	const double* pow_proximal;    // pow(max(1, d) / 400000.0, -1.53), d < 400000
	double read_through_penalty;   // 1 + pow((fraction - 0.25) * 20, 2)
	float cutoff; u32 apply_filter;
};

struct evalue_fn {
	cand_state c; annot_view an; evalue_inputs in;
	ARB_HD void operator()(u32 k) const {
		const u32 g1 = c.gene1[k], g2 = c.gene2[k];
		const i32 bp1 = c.bp1[k], bp2 = c.bp2[k];
		const u32 d1 = c.dir1[k], d2 = c.dir2[k];
		const u32 reads = cand_support(c, k);
		const double a = 10000.0 / an.gene_exonic_length[g1] * hd_max(in.partner_count[g1] - 1, 1);
		const double b = 10000.0 / an.gene_exonic_length[g2] * hd_max(in.partner_count[g2] - 1, 1);
		const float max_partners = (float) (a > b ? a : b);
		const u32 ri = reads < in.n_read_table ? reads : in.n_read_table - 1;
		const double scale = (double) in.mapped_reads / 20000000.0 * in.pow_reads[ri];
		float e = (float) ((double) max_partners * (1.0 > scale ? 1.0 : scale));
		const i32 g1s = an.gene_start[g1], g1e = an.gene_end[g1], g2s = an.gene_start[g2], g2e = an.gene_end[g2];
		const bool intragenic = g1 == g2 || (bp1 >= g2s - 10000 && bp1 <= g2e + 10000 && bp2 >= g1s - 10000 && bp2 <= g1e + 10000);
		const bool read_through = c.contig1[k] == c.contig2[k] && bp2 - bp1 < 400000 && d1 == DOWNSTREAM && d2 == UPSTREAM;
		if (intragenic) {
			e = (float) ((double) e * (2.0 / (in.intragenic_duplications + in.intragenic_inversions)));
			if (d1 == UPSTREAM && d2 == DOWNSTREAM) e = (float) ((float) e * (float) in.intragenic_duplications); // float *= unsigned: float arithmetic
			else if (d1 == d2) e = (float) ((float) e * (float) in.intragenic_inversions);
			if (reads >= 1) {
				e = (float) ((double) e * in.pow_intragenic[ri]);
				const i32 sd = spliced_distance(an, c.contig1[k], bp1, bp2, g1);
				if (sd < 1000) {
					e = (float) ((double) e * in.pow_spliced1000[sd < 0 ? 0 : sd]);
					if (sd < 400) e = (float) ((double) e * in.pow_spliced400[sd < 0 ? 0 : sd]);
				}
			}
			const double penalty = (double) in.spliced_same_gene / 0.25 / (double) in.spliced_different_genes;
			e = (float) ((double) e * (1.0 > penalty ? 1.0 : penalty));
		} else if (reads >= 1) {
			e = (float) ((double) e * in.pow_intergenic[ri]);
			const i32 dist = bp2 - bp1;
			if (read_through) e = (float) ((double) e * in.pow_read_through[dist < 1 ? 0 : dist]);
			else if (c.contig1[k] == c.contig2[k] && dist < 400000) e = (float) ((double) e * in.pow_proximal[dist < 1 ? 0 : dist]);
		}
		e = (float) ((double) e * (4.0 / (in.spliced_breakpoints + in.exonic_breakpoints + in.intronic_breakpoints + in.exonic_intronic_breakpoints)));
		const u8 bits = c.bits[k];
		u32 factor;
		if (bits & (CB_SPLICED1 | CB_SPLICED2)) factor = in.spliced_breakpoints;
		else if ((bits & CB_EXONIC1) && (bits & CB_EXONIC2)) factor = hd_max(in.spliced_breakpoints, in.exonic_breakpoints);
		else if (!(bits & CB_EXONIC1) && !(bits & CB_EXONIC2)) factor = hd_max(in.spliced_breakpoints, in.intronic_breakpoints);
		else factor = hd_max(in.spliced_breakpoints, in.exonic_intronic_breakpoints);
		e = e * (float) factor;
		if (in.read_through_fraction > 0.25 && read_through) e = (float) ((double) e * in.read_through_penalty);
		c.evalue[k] = e;
	}
};

struct relative_support_fn { // filter_relative_support.cpp:209-223
	cand_state c; annot_view an; float cutoff;
	ARB_HD void operator()(u32 k) const {
		if (c.filter[k] != F_none) return;
		const u32 g1 = c.gene1[k], g2 = c.gene2[k];
		const i32 bp1 = c.bp1[k], bp2 = c.bp2[k];
		const bool intragenic = g1 == g2 || (bp1 >= an.gene_start[g2] - 10000 && bp1 <= an.gene_end[g2] + 10000 && bp2 >= an.gene_start[g1] - 10000 && bp2 <= an.gene_end[g1] + 10000);
		if (!(c.evalue[k] < cutoff && !(intragenic && c.split_reads1[k] + c.split_reads2[k] == 0))) c.filter[k] = F_relative_support;
	}
};

} // namespace arb
