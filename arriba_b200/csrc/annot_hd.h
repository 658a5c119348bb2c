// annot_hd.h -- host/device queries on the flattened annotation (annot_view).
//
// Behavioural contract (what the results must equal), by reference location:
//   region lookup / point+range queries .... annotation.t.hpp:55-100 (get_annotation_by_coordinate, +-2 bp neighbour merge)
//   splice-site test ....................... annotation.cpp:379-429   (is_breakpoint_spliced, MAX_SPLICE_SITE_DISTANCE = 2)
//   spliced distance ....................... annotation.cpp:570-619   (get_spliced_distance)
//   set algebra on sorted id lists ......... annotation.t.hpp:47-53   (combine_annotations)
// The reference keeps std::map<position, sorted-set<ptr>> per contig; here a contig's regions are a
// sorted array of region ends with CSR item lists, searched with a binary search.
#pragma once
#include "model.h"

namespace arb {

// first region r of the contig with region_end[r] >= pos; returns hi (one past the contig's last region) if none
ARB_HD u32 region_lower_bound(const i32* region_end, u32 lo, u32 hi, i32 pos) {
	while (lo < hi) {
		u32 mid = lo + ((hi - lo) >> 1);
		if (region_end[mid] < pos) lo = mid + 1; else hi = mid;
	}
	return lo;
}

// grid (optional, host side): per contig one entry per 4,096-base bin = first region whose end is >= the bin's first position, plus a final entry; it narrows
// the binary search from the whole contig to the regions of one bin
enum { REGION_GRID_SHIFT = 12 };
struct region_index_view { const u32* begin; const i32* end; const u32* off; const u32* items; u32 n_contigs; const u32* grid; const u32* grid_begin; };
ARB_HD region_index_view exon_index(const annot_view& a) { region_index_view v = {a.exon_region_begin, a.exon_region_end, a.exon_region_off, a.exon_region_items, a.n_contigs, a.exon_grid, a.exon_grid_begin}; return v; }
ARB_HD region_index_view gene_index(const annot_view& a) { region_index_view v = {a.gene_region_begin, a.gene_region_end, a.gene_region_off, a.gene_region_items, a.n_contigs, a.gene_grid, a.gene_grid_begin}; return v; }
// region_lower_bound over the contig's regions [lo, hi), through the grid where there is one
ARB_HD u32 region_find(const region_index_view& ix, u32 contig, u32 lo, u32 hi, i32 pos) {
	if (ix.grid) {
		const u32 g0 = ix.grid_begin[contig], bins = ix.grid_begin[contig + 1] - g0 - 1;
		u32 b = pos < 0 ? 0u : (u32) pos >> REGION_GRID_SHIFT;
		if (b >= bins) b = bins - 1;
		const u32 first = ix.grid[g0 + b], last = ix.grid[g0 + b + 1];
		if (first > lo) lo = first;
		if (last + 1 < hi) hi = last + 1;
	}
	return region_lower_bound(ix.end, lo, hi, pos);
}

// fixed-capacity sorted id set used inside kernels (gene sets are tiny; overflow is reported, never silently dropped)
template <int CAP> struct idset {
	u32 v[CAP]; u32 n; bool overflow;
	ARB_HD idset(): n(0), overflow(false) {}
	ARB_HD void clear() { n = 0; }
	ARB_HD void insert(u32 x) { // keep ascending, unique
		u32 p = 0;
		while (p < n && v[p] < x) ++p;
		if (p < n && v[p] == x) return;
		if (n >= (u32) CAP) { overflow = true; return; }
		for (u32 k = n; k > p; --k) v[k] = v[k - 1];
		v[p] = x; ++n;
	}
	ARB_HD void assign(const u32* src, u32 cnt) { n = 0; for (u32 k = 0; k < cnt; ++k) { if (n >= (u32) CAP) { overflow = true; return; } v[n++] = src[k]; } }
	ARB_HD bool contains(u32 x) const { for (u32 k = 0; k < n; ++k) if (v[k] == x) return true; return false; }
};

// |a ∩ b| > 0 on ascending lists
ARB_HD bool sets_intersect(const u32* a, u32 na, const u32* b, u32 nb) {
	u32 i = 0, j = 0;
	while (i < na && j < nb) { if (a[i] < b[j]) ++i; else if (b[j] < a[i]) ++j; else return true; }
	return false;
}

// out = a ∩ b; if empty and make_union: out = a ∪ b   (annotation.t.hpp:47-53)
template <int CAP> ARB_HD void combine_sets(const u32* a, u32 na, const u32* b, u32 nb, idset<CAP>& out, bool make_union) {
	out.clear();
	u32 i = 0, j = 0;
	while (i < na && j < nb) { if (a[i] < b[j]) ++i; else if (b[j] < a[i]) ++j; else { out.insert(a[i]); ++i; ++j; } }
	if (out.n == 0 && make_union) {
		for (i = 0; i < na; ++i) out.insert(a[i]);
		for (j = 0; j < nb; ++j) out.insert(b[j]);
	}
}

// items of all regions that a point or range query returns (annotation.t.hpp:55-100)
template <int CAP> ARB_HD void query_index(const region_index_view& ix, u32 contig, i32 start, i32 end, idset<CAP>& out) {
	out.clear();
	if (contig >= ix.n_contigs) return;
	const u32 lo = ix.begin[contig], hi = ix.begin[contig + 1];
	if (start == end) {
		u32 r = region_find(ix, contig, lo, hi, start);
		if (r < hi) out.assign(ix.items + ix.off[r], ix.off[r + 1] - ix.off[r]);
		return;
	}
	if (start > end) { i32 t = start; start = end; end = t; }
	idset<CAP> rs, re;
	u32 r = region_find(ix, contig, lo, hi, start);
	if (r < hi) {
		rs.assign(ix.items + ix.off[r], ix.off[r + 1] - ix.off[r]);
		if (ix.end[r] - start <= 2) { // the region ends within 2 bp of the start: also take the next region
			++r;
			if (r < hi) for (u32 k = ix.off[r]; k < ix.off[r + 1]; ++k) rs.insert(ix.items[k]);
		}
	}
	r = region_find(ix, contig, lo, hi, end);
	if (r < hi) re.assign(ix.items + ix.off[r], ix.off[r + 1] - ix.off[r]);
	if (r != lo && hi > lo) {
		--r;
		if (end - ix.end[r] <= 2) for (u32 k = ix.off[r]; k < ix.off[r + 1]; ++k) re.insert(ix.items[k]);
	}
	combine_sets(rs.v, rs.n, re.v, re.n, out, true);
	out.overflow = out.overflow || rs.overflow || re.overflow;
}

// does any exon of `gene` in region r qualify as a splice site next to `bp`?  (annotation.cpp:379-402)
ARB_HD bool region_has_splice_site(const annot_view& a, u32 r, u32 gene, u32 direction, i32 bp) {
	for (u32 k = a.exon_region_off[r]; k < a.exon_region_off[r + 1]; ++k) {
		u32 e = a.exon_region_items[k];
		if (a.exon_gene[e] != gene) continue;
		const u8 fl = a.exon_flags[e];
		const bool single = !(fl & EF_HAS_PREV) && !(fl & EF_HAS_NEXT) && a.exon_cds_start[e] != -1; // single-exon transcript with CDS
		if (direction == UPSTREAM) {
			if (hd_abs(a.exon_start[e] - bp) <= 2 && ((fl & EF_HAS_PREV) || single || a.exon_start[e] == a.exon_cds_start[e])) return true;
		} else {
			if (hd_abs(a.exon_end[e] - bp) <= 2 && ((fl & EF_HAS_NEXT) || single || a.exon_end[e] == a.exon_cds_end[e])) return true;
		}
	}
	return false;
}

// annotation.cpp:404-429
ARB_HD bool is_breakpoint_spliced(const annot_view& a, u32 gene, u32 direction, i32 bp) {
	const u32 contig = a.gene_contig[gene];
	if (contig >= a.n_contigs) return false;
	const u32 lo = a.exon_region_begin[contig], hi = a.exon_region_begin[contig + 1];
	if (lo == hi) return false;
	const u32 r = region_lower_bound(a.exon_region_end, lo, hi, bp);
	if (r < hi) {
		if (region_has_splice_site(a, r, gene, direction, bp)) return true;
		if (r + 1 < hi && region_has_splice_site(a, r + 1, gene, direction, bp)) return true;
	}
	if (r != lo && region_has_splice_site(a, r - 1, gene, direction, bp)) return true;
	return false;
}

// annotation.cpp:570-619
ARB_HD i32 spliced_distance(const annot_view& a, u32 contig, i32 p1, i32 p2, u32 gene) {
	if (p1 > p2) { i32 t = p1; p1 = p2; p2 = t; }
	if (contig >= a.n_contigs) return p2 - p1;
	const u32 lo = a.exon_region_begin[contig], hi = a.exon_region_begin[contig + 1];
	if (lo == hi) return p2 - p1;
	u32 r = region_lower_bound(a.exon_region_end, lo, hi, p1);
	i32 distance = 0;
	if (r < hi && a.exon_region_end[r] < p2) { distance += a.exon_region_end[r] - p1; p1 = a.exon_region_end[r]; }
	for (; r < hi && a.exon_region_end[r] < p2; ++r) {
		if (a.exon_region_end[r] < p1) continue;
		// among this gene's exons in the region: the one with the smallest (covered length / skipped length) ratio, first wins ties
		i32 best_start = -1, best_end = -1, best_skip = -1;
		for (u32 k = a.exon_region_off[r]; k < a.exon_region_off[r + 1]; ++k) {
			u32 e = a.exon_region_items[k];
			if (a.exon_gene[e] != gene || !(a.exon_flags[e] & EF_HAS_NEXT) || a.exon_next_start[e] > p2) continue;
			i32 es = hd_max(p1, a.exon_start[e]), ee = hd_min(p2, a.exon_end[e]);
			i32 skip = a.exon_next_start[e] - es + 1;
			if (best_start == -1 || 1.0 * (ee - es) / skip < 1.0 * (best_end - best_start) / best_skip) { best_start = es; best_end = ee; best_skip = skip; }
		}
		if (best_start != -1) { distance += best_end - best_start + 1; p1 = best_start + best_skip - 1; }
	}
	distance += p2 - p1;
	return distance;
}

// extent of the union of a gene set (annotation.cpp:557-567): -1/-1 for an empty set
ARB_HD void gene_set_extent(const annot_view& a, const u32* genes, u32 n, i32& start, i32& end) {
	start = -1; end = -1;
	for (u32 k = 0; k < n; ++k) {
		i32 s = a.gene_start[genes[k]], e = a.gene_end[genes[k]];
		if (start == -1 || start > s) start = s;
		if (end == -1 || end < e) end = e;
	}
}

} // namespace arb
