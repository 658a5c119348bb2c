// annotate_hd.h -- annotation of the resident fragments on the device: gene sets, exonic flag and transcribed strand per alignment, dummy genes for breakpoints
// outside annotated genes. One thread per fragment (the rules couple the alignments of a fragment); gene sets are tiny.
//
// Behavioural contract:
//   pass 1 ......... arriba.cpp:165-205 (exon-based annotation, gene-level fallback), annotation.cpp:431-503 (annotate_alignment: splice-site support picks
//                    among overlapping genes and settles an ambiguous strand), annotation.cpp:505-555 (annotate_alignments: mates must agree),
//                    read_chimeric_alignments.cpp:775-790 (assign_strands_from_strandedness)
//   dummy genes .... arriba.cpp:207-260 (one dummy gene per cluster of unannotated breakpoints: 10 kb, never across a known gene)
//   pass 2 ......... arriba.cpp:262-319 (unannotated breakpoints get the dummy genes; several dummy genes on one alignment collapse to the one that holds the
//                    breakpoint)
// The region queries stream the (ascending) item lists of the one or two regions at either end of the range instead of materialising them
// (annotation.t.hpp:55-100: intersection of the start set and the end set, their union when the intersection is empty), so the number of overlapping exon
// records does not matter; only the resulting GENE sets are held (idset<CAP>, overflow is reported through `error`, never silently dropped).
#pragma once
#include "model.h"
#include "annot_hd.h"
#include "prims.h"

namespace arb {

// ascending union of up to two ascending id lists
struct merged2 {
	const u32* a; u32 na; const u32* b; u32 nb; u32 i, j;
	ARB_HD bool done() const { return i >= na && j >= nb; }
	ARB_HD u32 peek() const { if (i >= na) return b[j]; if (j >= nb) return a[i]; return a[i] < b[j] ? a[i] : b[j]; }
	ARB_HD void pop() { const u32 v = peek(); if (i < na && a[i] == v) ++i; if (j < nb && b[j] == v) ++j; }
};

// calls emit(id) for the ids a point (start == end) or range query returns, ascending (same result set as query_index, annot_hd.h)
template <class E> ARB_HD void stream_query(const region_index_view& ix, u32 contig, i32 start, i32 end, E& emit) {
	if (contig >= ix.n_contigs) return;
	const u32 lo = ix.begin[contig], hi = ix.begin[contig + 1];
	if (start == end) {
		const u32 r = region_find(ix, contig, lo, hi, start);
		if (r < hi) for (u32 k = ix.off[r]; k < ix.off[r + 1]; ++k) emit(ix.items[k]);
		return;
	}
	if (start > end) { const i32 t = start; start = end; end = t; }
	merged2 rs = {0, 0, 0, 0, 0, 0}, re = {0, 0, 0, 0, 0, 0};
	u32 r = region_find(ix, contig, lo, hi, start);
	if (r < hi) {
		rs.a = ix.items + ix.off[r]; rs.na = ix.off[r + 1] - ix.off[r];
		if (ix.end[r] - start <= 2) { ++r; if (r < hi) { rs.b = ix.items + ix.off[r]; rs.nb = ix.off[r + 1] - ix.off[r]; } }
	}
	r = region_find(ix, contig, lo, hi, end);
	if (r < hi) { re.a = ix.items + ix.off[r]; re.na = ix.off[r + 1] - ix.off[r]; }
	if (r != lo && hi > lo) { --r; if (end - ix.end[r] <= 2) { re.b = ix.items + ix.off[r]; re.nb = ix.off[r + 1] - ix.off[r]; } }
	merged2 x = rs, y = re; u32 common = 0;
	while (!x.done() && !y.done()) {
		const u32 xv = x.peek(), yv = y.peek();
		if (xv < yv) x.pop(); else if (yv < xv) y.pop(); else { emit(xv); ++common; x.pop(); y.pop(); }
	}
	if (common) return;
	x = rs; y = re;
	while (!x.done() || !y.done()) {
		const u32 v = x.done() ? y.peek() : y.done() ? x.peek() : hd_min(x.peek(), y.peek());
		emit(v);
		if (!x.done() && x.peek() == v) x.pop();
		if (!y.done() && y.peek() == v) y.pop();
	}
}

template <int CAP> struct emit_exon_gene { const u32* exon_gene; idset<CAP>* out; ARB_HD void operator()(u32 e) { out->insert(exon_gene[e]); } };
template <int CAP> struct emit_id { idset<CAP>* out; ARB_HD void operator()(u32 g) { out->insert(g); } };
template <int CAP> ARB_HD void genes_by_exons(const annot_view& an, u32 contig, i32 start, i32 end, idset<CAP>& out) { out.clear(); emit_exon_gene<CAP> e = {an.exon_gene, &out}; stream_query(exon_index(an), contig, start, end, e); }
template <int CAP> ARB_HD void genes_by_position(const annot_view& an, u32 contig, i32 start, i32 end, idset<CAP>& out) { out.clear(); emit_id<CAP> e = {&out}; stream_query(gene_index(an), contig, start, end, e); }

// gene sets between the passes: up to ROW ids in place, larger sets in a pool (bump allocation; the final CSR is laid out afterwards in alignment order)
struct gene_sets_view {
	enum { ROW = 4 };
	u32* rows; u16* cnt; u32* pool; u32* pool_top; u32 pool_cap; u32* error; // error bits: 1 gene set overflow, 2 pool exhausted
	ARB_HD const u32* get(size_t a) const { return cnt[a] <= ROW ? rows + a * ROW : pool + rows[a * ROW]; }
	ARB_HD void set(size_t a, const u32* g, u32 n) const {
		cnt[a] = (u16) n;
		u32* dst = rows + a * ROW;
		if (n > ROW) {
			const u32 at = atomic_add_u32(pool_top, n);
			if ((u64) at + n > pool_cap) { atomic_or_u32(error, 2u); cnt[a] = 0; return; }
			rows[a * ROW] = at; dst = pool + at;
		}
		for (u32 k = 0; k < n; ++k) dst[k] = g[k];
	}
	template <int CAP> ARB_HD void load(size_t a, idset<CAP>& s) const { s.clear(); s.assign(get(a), cnt[a]); }
};

ARB_HD bool pred_amb(const frag_view& f, u32 a) { return f.aflags[a] & AF_PRED_AMBIGUOUS; }
ARB_HD bool pred_fwd(const frag_view& f, u32 a) { return f.aflags[a] & AF_PRED_FORWARD; }
ARB_HD void set_pred(const frag_view& f, u32 a, bool forward) { f.aflags[a] = (u8) ((f.aflags[a] & ~(AF_PRED_AMBIGUOUS | AF_PRED_FORWARD)) | (forward ? AF_PRED_FORWARD : 0)); }
ARB_HD void set_amb(const frag_view& f, u32 a) { f.aflags[a] |= AF_PRED_AMBIGUOUS; }

// gene set and strand of one alignment from the exon index (annotation.cpp:431-503)
template <int CAP> ARB_HD void annotate_alignment(const annot_view& an, const frag_view& f, u32 a, idset<CAP>& genes) {
	genes_by_exons(an, f.contig[a], f.start[a], f.end[a], genes);
	const bool ambiguous_strand = f.aflags[a] & AF_PRED_AMBIGUOUS;
	if (!(f.cigar_cnt[a] > 1 && (genes.n > 1 || ambiguous_strand))) return;
	// look for a clip or intron whose position coincides with a splice site of only some of the genes
	idset<CAP> supported;
	i32 ref = f.start[a];
	const u32* c = f.cig(a);
	for (u32 i = 0; i < f.cigar_cnt[a] && supported.n == 0; ++i) {
		const u32 op = cig_op(c[i]); const i32 len = (i32) cig_len(c[i]);
		if (op == C_S || op == C_H || op == C_N) {
			supported.clear();
			for (u32 k = 0; k < genes.n; ++k) {
				const u32 g = genes.v[k];
				bool drop;
				if (op == C_N) drop = !is_breakpoint_spliced(an, g, DOWNSTREAM, ref) && !is_breakpoint_spliced(an, g, UPSTREAM, ref + len);
				else drop = i == 0 ? !is_breakpoint_spliced(an, g, UPSTREAM, ref) : !is_breakpoint_spliced(an, g, DOWNSTREAM, ref);
				if (!drop) supported.insert(g);
			}
		}
		if (op == C_N || op == C_M || op == C_X || op == C_EQ || op == C_D) ref += len;
	}
	if (supported.n == 0) return;
	if (supported.n < genes.n) { genes.n = supported.n; for (u32 k = 0; k < supported.n; ++k) genes.v[k] = supported.v[k]; }
	if (ambiguous_strand) {
		const u8 strand = an.gene_strand[supported.v[0]];
		bool consistent = true;
		for (u32 k = 0; k < supported.n; ++k) if (an.gene_strand[supported.v[k]] != strand) consistent = false;
		if (consistent) f.aflags[a] = (u8) ((f.aflags[a] & ~(AF_PRED_AMBIGUOUS | AF_PRED_FORWARD)) | (strand ? AF_PRED_FORWARD : 0));
	}
}

template <int CAP> ARB_HD void narrow_to(idset<CAP>& g, const idset<CAP>& combined) { if (g.n == 0 || combined.n < g.n) { g.n = combined.n; for (u32 k = 0; k < combined.n; ++k) g.v[k] = combined.v[k]; } }

// pass 1 of one fragment
template <int CAP> struct annotate_pass1_fn {
	annot_view an; frag_view f; gene_sets_view sets; int strandedness;
	ARB_HD void operator()(u32 i) const {
		const u32 na = f.n_aln[i];
		const u32 a[3] = {f.idx(i, 0), f.idx(i, 1), f.idx(i, 2)};
		if (strandedness != 0) { // read_chimeric_alignments.cpp:775-790
			const bool first_is_mate1 = f.aflags[a[MATE1]] & AF_FIRST_IN_PAIR;
			const u32 first = first_is_mate1 ? a[MATE1] : a[MATE2], second = first_is_mate1 ? a[MATE2] : a[MATE1];
			const bool ps_first = (strandedness == 2) ? !f.fwd(first) : f.fwd(first);
			set_pred(f, first, ps_first);
			set_pred(f, second, f.fwd(first) == f.fwd(second) ? !ps_first : ps_first);
			if (na == 3) { const u32 s = a[SPLIT_READ], u = a[SUPPLEMENTARY]; set_pred(f, u, f.fwd(u) != f.fwd(s) ? !pred_fwd(f, s) : pred_fwd(f, s)); }
		}
		idset<CAP> g[3], combined;
		for (u32 s = 0; s < na; ++s) {
			annotate_alignment(an, f, a[s], g[s]);
			if (g[s].n) f.aflags[a[s]] |= AF_EXONIC; else f.aflags[a[s]] &= (u8) ~AF_EXONIC;
		}
		// strands of the two mates must be consistent (annotation.cpp:514-526)
		if (pred_amb(f, a[0]) && !pred_amb(f, a[1])) set_pred(f, a[0], f.fwd(a[0]) == f.fwd(a[1]) ? !pred_fwd(f, a[1]) : pred_fwd(f, a[1]));
		else if (!pred_amb(f, a[0]) && pred_amb(f, a[1])) set_pred(f, a[1], f.fwd(a[0]) == f.fwd(a[1]) ? !pred_fwd(f, a[0]) : pred_fwd(f, a[0]));
		else if (!pred_amb(f, a[0]) && !pred_amb(f, a[1])) {
			if ((pred_fwd(f, a[0]) != pred_fwd(f, a[1])) != (f.fwd(a[0]) == f.fwd(a[1]))) { set_amb(f, a[0]); set_amb(f, a[1]); }
		}
		if (na == 3) {
			combine_sets(g[1].v, g[1].n, g[0].v, g[0].n, combined, true);
			narrow_to(g[0], combined); narrow_to(g[1], combined);
			const bool differ = f.fwd(a[2]) != f.fwd(a[1]);
			if (pred_amb(f, a[1]) && !pred_amb(f, a[2])) { const bool ps = differ ? !pred_fwd(f, a[2]) : pred_fwd(f, a[2]); set_pred(f, a[0], ps); set_pred(f, a[1], ps); }
			else if (!pred_amb(f, a[1]) && pred_amb(f, a[2])) set_pred(f, a[2], differ ? !pred_fwd(f, a[1]) : pred_fwd(f, a[1]));
			else if (!pred_amb(f, a[1]) && !pred_amb(f, a[2])) {
				if ((pred_fwd(f, a[1]) != pred_fwd(f, a[2])) != differ) { set_amb(f, a[0]); set_amb(f, a[1]); set_amb(f, a[2]); }
			}
		}
		// gene-level fallback for alignments that hit no exon
		for (u32 s = 0; s < na; ++s) if (g[s].n == 0) genes_by_position(an, f.contig[a[s]], f.start[a[s]], f.end[a[s]], g[s]);
		if (na == 3) {
			combine_sets(g[1].v, g[1].n, g[0].v, g[0].n, combined, true);
			narrow_to(g[0], combined); narrow_to(g[1], combined);
		}
		bool overflow = combined.overflow;
		for (u32 s = 0; s < na; ++s) { overflow = overflow || g[s].overflow; sets.set(a[s], g[s].v, g[s].n); }
		for (u32 s = na; s < 3; ++s) sets.cnt[a[s]] = 0;
		if (overflow) atomic_or_u32(sets.error, 1u);
	}
};

// breakpoints without a gene after pass 1 (arriba.cpp:214-232): up to two per fragment, as (contig, position)
ARB_HD u32 unmapped_breakpoints(const frag_view& f, const u16* cnt, u32 i, u32* contig, u32* pos) {
	u32 n = 0;
	if (f.n_aln[i] == 3) {
		const u32 s = f.idx(i, SPLIT_READ), u = f.idx(i, SUPPLEMENTARY);
		if (cnt[s] == 0) { if (contig) { contig[n] = f.contig[s]; pos[n] = (u32) (f.fwd(s) ? f.start[s] : f.end[s]); } ++n; }
		if (cnt[u] == 0) { if (contig) { contig[n] = f.contig[u]; pos[n] = (u32) (f.fwd(u) ? f.end[u] : f.start[u]); } ++n; }
	} else for (u32 s = 0; s < 2; ++s) {
		const u32 a = f.idx(i, s);
		if (cnt[a] == 0) { if (contig) { contig[n] = f.contig[a]; pos[n] = (u32) (f.fwd(a) ? f.end[a] : f.start[a]); } ++n; }
	}
	return n;
}
struct unmapped_count_fn { frag_view f; const u16* cnt; u32* out; ARB_HD void operator()(u32 i) const { out[i] = unmapped_breakpoints(f, cnt, i, NULL, NULL); } };
struct unmapped_fill_fn {
	frag_view f; const u16* cnt; const u32* off; u32* contig; u32* pos;
	ARB_HD void operator()(u32 i) const { u32 c[2], p[2]; const u32 n = unmapped_breakpoints(f, cnt, i, c, p); for (u32 k = 0; k < n; ++k) { contig[off[i] + k] = c[k]; pos[off[i] + k] = p[k]; } }
};
// The reference sweeps the sorted breakpoints and closes a cluster when the next one is more than 10 kb past the last, lies on another contig, or has reached
// the end of the first known-gene region at or after the cluster's START (arriba.cpp:236-256). While a cluster stays open that region is also the first one
// at or after the PREVIOUS breakpoint (every breakpoint since the start lies before its end), so the test only needs the two neighbours.
struct dummy_break_fn {
	annot_view an; const u32* contig; const u32* pos; u32* brk;
	ARB_HD void operator()(u32 k) const {
		if (k == 0) { brk[k] = 1; return; }
		const u32 c = contig[k - 1]; const i32 prev = (i32) pos[k - 1], cur = (i32) pos[k];
		bool b = contig[k] != c || prev + 10000 < cur;
		if (!b && c < an.n_contigs) {
			const u32 lo = an.gene_region_begin[c], hi = an.gene_region_begin[c + 1];
			const u32 r = region_lower_bound(an.gene_region_end, lo, hi, prev);
			b = r != hi && an.gene_region_end[r] <= cur;
		}
		brk[k] = b ? 1u : 0u;
	}
};
struct dummy_emit_fn {
	const u32* contig; const u32* pos; const u32* brk; const u32* brk_scan; u32 n; u16* d_contig; i32* d_start; i32* d_end;
	ARB_HD void operator()(u32 k) const {
		const u32 id = brk_scan[k] + brk[k] - 1;
		if (brk[k]) { d_contig[id] = (u16) contig[k]; d_start[id] = (i32) pos[k]; }
		if (k + 1 == n || brk[k + 1]) d_end[id] = (i32) pos[k];
	}
};

// pass 2 of one fragment (arriba.cpp:262-319); `an` now holds the dummy genes. Gene sets are handled as views (pointer, size): a set of pass 1, the item
// list of the region a breakpoint lies in (the result of a point query, annotation.t.hpp:55-68), or a single picked gene -- nothing is materialised, so a
// breakpoint covered by hundreds of stacked dummy genes (many reads clipped exactly at the end of a gene region, each the start of a cluster of its own,
// arriba.cpp:236-256) costs no memory.
struct gene_list { const u32* v; u32 n; };
ARB_HD gene_list genes_at(const annot_view& an, u32 contig, i32 pos) {
	gene_list g = {0, 0};
	if (contig >= an.n_contigs) return g;
	const u32 lo = an.gene_region_begin[contig], hi = an.gene_region_begin[contig + 1];
	const u32 r = region_lower_bound(an.gene_region_end, lo, hi, pos);
	if (r < hi) { g.v = an.gene_region_items + an.gene_region_off[r]; g.n = an.gene_region_off[r + 1] - an.gene_region_off[r]; }
	return g;
}
struct annotate_pass2_fn {
	annot_view an; frag_view f; gene_sets_view sets;
	ARB_HD void operator()(u32 i) const {
		const u32 na = f.n_aln[i];
		const u32 a[3] = {f.idx(i, 0), f.idx(i, 1), f.idx(i, 2)};
		bool changed[3] = {false, false, false};
		gene_list g[3] = {{0, 0}, {0, 0}, {0, 0}};
		u32 single[3]; // a set collapsed to one gene
		for (u32 s = 0; s < na; ++s) { g[s].v = sets.get(a[s]); g[s].n = sets.cnt[a[s]]; }
		if (na == 3) {
			if (g[0].n == 0 || g[1].n == 0) {
				const i32 bp = f.fwd(a[1]) ? f.start[a[1]] : f.end[a[1]];
				g[1] = genes_at(an, f.contig[a[1]], bp); g[0] = g[1];
				changed[0] = changed[1] = true;
			}
			if (g[2].n == 0) { const i32 bp = f.fwd(a[2]) ? f.end[a[2]] : f.start[a[2]]; g[2] = genes_at(an, f.contig[a[2]], bp); changed[2] = true; }
		} else {
			for (u32 s = 0; s < 2; ++s) if (g[s].n == 0) { const i32 bp = f.fwd(a[s]) ? f.end[a[s]] : f.start[a[s]]; g[s] = genes_at(an, f.contig[a[s]], bp); changed[s] = true; }
		}
		// several dummy genes on one alignment: keep the one that contains the breakpoint (default: MATE1's first gene)
		const u32 mate1_first = g[0].n ? g[0].v[0] : 0;
		for (u32 s = 0; s < na; ++s) {
			if (g[s].n > 1 && (an.gene_flags[g[s].v[0]] & GF_DUMMY)) {
				const i32 bp = f.fwd(a[s]) ? f.start[a[s]] : f.end[a[s]];
				u32 pick = s == 0 ? g[0].v[0] : (g[0].n ? g[0].v[0] : mate1_first);
				for (u32 k = 0; k < g[s].n; ++k) if (an.gene_start[g[s].v[k]] <= bp && an.gene_end[g[s].v[k]] >= bp) pick = g[s].v[k];
				single[s] = pick; g[s].v = &single[s]; g[s].n = 1; changed[s] = true;
			}
		}
		if (na == 3 && g[0].n && g[1].n && g[0].v[0] != g[1].v[0] && (an.gene_flags[g[0].v[0]] & GF_DUMMY) && (an.gene_flags[g[1].v[0]] & GF_DUMMY)) {
			const i32 bp = f.fwd(a[1]) ? f.start[a[1]] : f.end[a[1]];
			u32 pick = g[0].v[0];
			for (u32 k = 0; k < g[0].n; ++k) if (an.gene_start[g[0].v[k]] <= bp && an.gene_end[g[0].v[k]] >= bp) pick = g[0].v[k];
			for (u32 k = 0; k < g[1].n; ++k) if (an.gene_start[g[1].v[k]] <= bp && an.gene_end[g[1].v[k]] >= bp) pick = g[1].v[k];
			single[0] = pick; single[1] = pick; g[0].v = &single[0]; g[0].n = 1; g[1].v = &single[1]; g[1].n = 1; changed[0] = changed[1] = true;
		}
		for (u32 s = 0; s < na; ++s) if (changed[s]) {
			if (g[s].n > 0xFFFFu) { atomic_or_u32(sets.error, 1u); continue; } // the gene count of an alignment is a 16-bit column
			sets.set(a[s], g[s].v, g[s].n);
		}
	}
};

// final CSR gene columns in alignment order
struct gene_count_fn { const u16* cnt; u32* out; ARB_HD void operator()(u32 a) const { out[a] = cnt[a]; } };
struct gene_fill_fn {
	gene_sets_view sets; const u32* off; frag_view f;
	ARB_HD void operator()(u32 a) const {
		const u32 n = sets.cnt[a]; const u32* src = sets.get(a);
		f.genes_off[a] = off[a]; f.genes_cnt[a] = (u16) n;
		for (u32 k = 0; k < n; ++k) f.genes[off[a] + k] = src[k];
	}
};

} // namespace arb
