// annotate.cpp -- annotation of fragments with genes / exonic flag / transcribed strand, creation of dummy genes for
// intergenic breakpoints, strandedness detection and the fragment-length statistics.
// Behavioural contract: arriba.cpp:150-325 (inline annotation passes), annotation.cpp:431-555 (annotate_alignment(s)),
// read_stats.cpp:11-146 (estimate_fragment_length, detect_strandedness), read_chimeric_alignments.cpp:775-790
// (assign_strands_from_strandedness). Gene sets are ordered by gene id (creation order), see refdata.h.
#include "pipeline.h"
#include "../annot_hd.h"
#include "index_query.h"
#include <algorithm>
#include <cmath>
#include <iostream>
#include <list>
#include <thread>
#include <cstring>

namespace arb { namespace host {

typedef idset<1024> gset;       // gene set of one alignment (an alignment with a long intron gap spans many genes)
enum { EXON_SET_CAPACITY = 16384 }; // exons under one alignment (reached through index_query: a small set first)

template <class F> static void parallel_ranges(int threads, size_t n, F f) {
	if (threads <= 1 || n < 1024) { f(0, (size_t) 0, n); return; }
	std::vector<std::thread> pool; std::vector<std::string> errors(threads);
	for (int t = 0; t < threads; ++t) pool.emplace_back([&, t]() { try { f(t, n * t / threads, n * (t + 1) / threads); } catch (const std::exception& e) { errors[t] = e.what(); } });
	for (size_t t = 0; t < pool.size(); ++t) pool[t].join();
	for (int t = 0; t < threads; ++t) if (!errors[t].empty()) throw std::runtime_error(errors[t]);
}

// gene sets under construction: per alignment a reference into one of the per-thread pools
struct gene_refs {
	std::vector<std::vector<u32> > pools; column<u64> off; std::vector<u16> cnt; column<u8> pool;
	void init(size_t n_aln, int threads) { pools.assign(threads, std::vector<u32>()); off.resize(n_aln); cnt.assign(n_aln, 0); pool.resize(n_aln); } // only `cnt` has to start at zero
	const u32* get(size_t a) const { return cnt[a] ? pools[pool[a]].data() + off[a] : NULL; }
	void set(size_t a, int thread, const u32* g, u32 n) { std::vector<u32>& p = pools[thread]; off[a] = p.size(); cnt[a] = (u16) n; pool[a] = (u8) thread; p.insert(p.end(), g, g + n); }
	void load(size_t a, gset& s) const { s.clear(); s.assign(get(a), cnt[a]); }
};

template <int CAP> static void check_overflow(const idset<CAP>& s) { if (s.overflow) throw std::runtime_error("more than " + std::to_string(CAP) + " overlapping genes or exons under one alignment are not supported"); }

// gene set and strand of one alignment from the exon index (annotation.cpp:431-503)
static void annotate_alignment(const annot_view& an, const frag_view& f, u32 a, gset& genes) {
	genes.clear();
	index_query<EXON_SET_CAPACITY>(exon_index(an), f.contig[a], f.start[a], f.end[a], [&](const u32* exons_hit, u32 n) { for (u32 k = 0; k < n; ++k) genes.insert(an.exon_gene[exons_hit[k]]); },
	                               "more than 16384 overlapping genes or exons under one alignment are not supported");
	check_overflow(genes);
	const bool ambiguous_strand = f.aflags[a] & AF_PRED_AMBIGUOUS;
	if (!(f.cigar_cnt[a] > 1 && (genes.n > 1 || ambiguous_strand))) return;
	// look for a clip or intron whose position coincides with a splice site of only some of the genes
	gset supported;
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
	if (supported.n < genes.n) genes = supported;
	if (ambiguous_strand) {
		const u8 strand = an.gene_strand[supported.v[0]];
		bool consistent = true;
		for (u32 k = 0; k < supported.n; ++k) if (an.gene_strand[supported.v[k]] != strand) consistent = false;
		if (consistent) f.aflags[a] = (u8) ((f.aflags[a] & ~(AF_PRED_AMBIGUOUS | AF_PRED_FORWARD)) | (strand ? AF_PRED_FORWARD : 0));
	}
}

static inline bool pred_amb(const frag_view& f, u32 a) { return f.aflags[a] & AF_PRED_AMBIGUOUS; }
static inline bool pred_fwd(const frag_view& f, u32 a) { return f.aflags[a] & AF_PRED_FORWARD; }
static inline void set_pred(const frag_view& f, u32 a, bool forward) { f.aflags[a] = (u8) ((f.aflags[a] & ~(AF_PRED_AMBIGUOUS | AF_PRED_FORWARD)) | (forward ? AF_PRED_FORWARD : 0)); }
static inline void set_amb(const frag_view& f, u32 a) { f.aflags[a] |= AF_PRED_AMBIGUOUS; }

// ------------------------------------------------------------------------------------------- strandedness
int detect_strandedness(pipeline& p) { // read_stats.cpp:94-146; returns 0 no, 1 yes, 2 reverse
	const annot_view an = p.ref.host_view(); const frag_view f = p.frags.view();
	u32 count = 0, matching = 0;
	gset genes;
	for (u32 i = 0; i < f.n && count < 100; ++i) {
		if (f.n_aln[i] != 3) continue;
		const u32 m = f.idx(i, MATE1), s = f.idx(i, SPLIT_READ), u = f.idx(i, SUPPLEMENTARY);
		if (!(f.contig[s] == f.contig[u] && f.fwd(s) == f.fwd(u) && std::abs(f.start[s] - f.start[u]) < 400000)) continue;
		query_index(gene_index(an), f.contig[s], f.start[s], f.end[s], genes); check_overflow(genes);
		if (genes.n != 1) continue;
		const u32 g = genes.v[0];
		if (!is_breakpoint_spliced(an, g, f.fwd(s) ? UPSTREAM : DOWNSTREAM, f.fwd(s) ? f.start[s] : f.end[s])) continue;
		const bool gene_fwd = an.gene_strand[g];
		if (((f.aflags[s] & AF_FIRST_IN_PAIR) && f.fwd(s) == gene_fwd) || ((f.aflags[m] & AF_FIRST_IN_PAIR) && f.fwd(m) == gene_fwd)) ++matching;
		++count;
	}
	if (count < 100) return 0;
	if (matching < (1 - 0.95f) * count) return 2;
	if (matching > 0.95f * count) return 1;
	return 0;
}

void assign_strands(pipeline& p, int strandedness) { // read_chimeric_alignments.cpp:775-790
	if (strandedness == 0) return;
	const frag_view f = p.frags.view();
	parallel_ranges(p.threads, f.n, [&](int, size_t lo, size_t hi) {
		for (size_t i = lo; i < hi; ++i) {
			const bool first_is_mate1 = f.aflags[f.idx((u32) i, MATE1)] & AF_FIRST_IN_PAIR;
			const u32 first = f.idx((u32) i, first_is_mate1 ? MATE1 : MATE2), second = f.idx((u32) i, first_is_mate1 ? MATE2 : MATE1);
			const bool ps_first = (strandedness == 2) ? !f.fwd(first) : f.fwd(first);
			set_pred(f, first, ps_first);
			set_pred(f, second, f.fwd(first) == f.fwd(second) ? !ps_first : ps_first);
			if (f.n_aln[i] == 3) {
				const u32 s = f.idx((u32) i, SPLIT_READ), u = f.idx((u32) i, SUPPLEMENTARY);
				set_pred(f, u, f.fwd(u) != f.fwd(s) ? !pred_fwd(f, s) : pred_fwd(f, s));
			}
		}
	});
}

// ------------------------------------------------------------------------------------------- annotation passes
void annotate_fragments(pipeline& p) {
	refdata& ref = p.ref;
	fragment_table& ft = p.frags;
	const u32 n = ft.n;
	const int T = std::max(1, p.threads);
	stage_laps laps("annotate");
	gene_refs refs; refs.init(3 * (size_t) n, T);
	laps.lap("reference table");
	{
		const annot_view an = ref.host_view(); const frag_view f = ft.view();
		// pass 1: exon-based annotation, mate-aware strand inference, gene-level fallback (arriba.cpp:186-205)
		parallel_ranges(T, n, [&](int t, size_t lo, size_t hi) {
			gset g[3], combined;
			for (size_t i = lo; i < hi; ++i) {
				const u32 na = f.n_aln[i];
				u32 a[3] = {f.idx((u32) i, 0), f.idx((u32) i, 1), f.idx((u32) i, 2)};
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
					if (g[0].n == 0 || combined.n < g[0].n) g[0] = combined;
					if (g[1].n == 0 || combined.n < g[1].n) g[1] = combined;
					const bool differ = f.fwd(a[2]) != f.fwd(a[1]);
					if (pred_amb(f, a[1]) && !pred_amb(f, a[2])) { const bool ps = differ ? !pred_fwd(f, a[2]) : pred_fwd(f, a[2]); set_pred(f, a[0], ps); set_pred(f, a[1], ps); }
					else if (!pred_amb(f, a[1]) && pred_amb(f, a[2])) set_pred(f, a[2], differ ? !pred_fwd(f, a[1]) : pred_fwd(f, a[1]));
					else if (!pred_amb(f, a[1]) && !pred_amb(f, a[2])) {
						if ((pred_fwd(f, a[1]) != pred_fwd(f, a[2])) != differ) { set_amb(f, a[0]); set_amb(f, a[1]); set_amb(f, a[2]); }
					}
				}
				// gene-level fallback for alignments that hit no exon
				for (u32 s = 0; s < na; ++s) if (g[s].n == 0) { query_index(gene_index(an), f.contig[a[s]], f.start[a[s]], f.end[a[s]], g[s]); check_overflow(g[s]); }
				if (na == 3) {
					combine_sets(g[1].v, g[1].n, g[0].v, g[0].n, combined, true);
					if (g[0].n == 0 || combined.n < g[0].n) g[0] = combined;
					if (g[1].n == 0 || combined.n < g[1].n) g[1] = combined;
				}
				for (u32 s = 0; s < na; ++s) { check_overflow(g[s]); refs.set(a[s], t, g[s].v, g[s].n); }
			}
		});
	}

	laps.lap("pass 1 (exons, strands)");
	// dummy genes for breakpoints outside annotated genes: one dummy gene per 10 kb cluster (arriba.cpp:207-260)
	struct unmapped_t { u16 contig; i32 pos; };
	std::vector<unmapped_t> unmapped;
	{
		const frag_view f = ft.view();
		std::vector<std::vector<unmapped_t> > part(T);
		parallel_ranges(T, n, [&](int t, size_t lo, size_t hi) {
			std::vector<unmapped_t>& out = part[t];
			for (size_t k = lo; k < hi; ++k) {
				const u32 i = (u32) k;
				if (f.n_aln[i] == 3) {
					const u32 s = f.idx(i, SPLIT_READ), u = f.idx(i, SUPPLEMENTARY);
					if (refs.cnt[s] == 0) { unmapped_t x = {f.contig[s], f.fwd(s) ? f.start[s] : f.end[s]}; out.push_back(x); }
					if (refs.cnt[u] == 0) { unmapped_t x = {f.contig[u], f.fwd(u) ? f.end[u] : f.start[u]}; out.push_back(x); }
				} else for (u32 s = 0; s < 2; ++s) {
					const u32 a = f.idx(i, s);
					if (refs.cnt[a] == 0) { unmapped_t x = {f.contig[a], f.fwd(a) ? f.end[a] : f.start[a]}; out.push_back(x); }
				}
			}
		});
		for (int t = 0; t < T; ++t) unmapped.insert(unmapped.end(), part[t].begin(), part[t].end());
	}
	const size_t first_dummy = ref.genes.size();
	if (!unmapped.empty()) {
		std::stable_sort(unmapped.begin(), unmapped.end(), [](const unmapped_t& a, const unmapped_t& b) { return a.contig != b.contig ? a.contig < b.contig : a.pos < b.pos; });
		const region_index& gi = ref.gene_index;
		auto next_region = [&](u16 contig, i32 pos) { const u32 lo = gi.begin[contig], hi = gi.begin[contig + 1]; return (u32) (std::lower_bound(gi.end.begin() + lo, gi.end.begin() + hi, pos) - gi.end.begin()); };
		gene_rec d; d.forward = true; d.exonic_length = 10000; d.is_dummy = true; d.is_protein_coding = false;
		d.contig = unmapped[0].contig; d.start = unmapped[0].pos; d.end = unmapped[0].pos;
		u32 next_known = next_region(d.contig, unmapped[0].pos);
		for (size_t k = 1;; ++k) {
			const bool at_end = k == unmapped.size();
			if (at_end || d.end + 10000 < unmapped[k].pos || (next_known != gi.begin[d.contig + 1] && gi.end[next_known] <= unmapped[k].pos) || unmapped[k].contig != d.contig) {
				ref.genes.push_back(d);
				if (at_end) break;
				d.contig = unmapped[k].contig; d.start = unmapped[k].pos;
				next_known = next_region(d.contig, unmapped[k].pos);
			}
			d.end = unmapped[k].pos;
		}
	}
	(void) first_dummy;
	laps.lap("dummy genes");
	ref.build_gene_index();
	ref.flatten();
	laps.lap("gene index rebuilt");

	// pass 2: map still unannotated breakpoints to the dummy genes, then collapse multi-dummy annotations (arriba.cpp:262-319)
	{
		const annot_view an = ref.host_view(); const frag_view f = ft.view();
		parallel_ranges(T, n, [&](int t, size_t lo, size_t hi) {
			gset g[3];
			for (size_t i = lo; i < hi; ++i) {
				const u32 na = f.n_aln[i];
				u32 a[3] = {f.idx((u32) i, 0), f.idx((u32) i, 1), f.idx((u32) i, 2)};
				bool changed[3] = {false, false, false};
				for (u32 s = 0; s < na; ++s) refs.load(a[s], g[s]);
				if (na == 3) {
					if (g[0].n == 0 || g[1].n == 0) {
						const i32 bp = f.fwd(a[1]) ? f.start[a[1]] : f.end[a[1]];
						query_index(gene_index(an), f.contig[a[1]], bp, bp, g[1]); g[0] = g[1]; changed[0] = changed[1] = true;
					}
					if (g[2].n == 0) { const i32 bp = f.fwd(a[2]) ? f.end[a[2]] : f.start[a[2]]; query_index(gene_index(an), f.contig[a[2]], bp, bp, g[2]); changed[2] = true; }
				} else {
					for (u32 s = 0; s < 2; ++s) if (g[s].n == 0) { const i32 bp = f.fwd(a[s]) ? f.end[a[s]] : f.start[a[s]]; query_index(gene_index(an), f.contig[a[s]], bp, bp, g[s]); changed[s] = true; }
				}
				// several dummy genes on one alignment: keep the one that contains the breakpoint (default: MATE1's first gene)
				const u32 mate1_first = g[0].n ? g[0].v[0] : 0;
				for (u32 s = 0; s < na; ++s) {
					if (g[s].n > 1 && (an.gene_flags[g[s].v[0]] & GF_DUMMY)) {
						const i32 bp = f.fwd(a[s]) ? f.start[a[s]] : f.end[a[s]];
						u32 pick = s == 0 ? g[0].v[0] : (g[0].n ? g[0].v[0] : mate1_first);
						for (u32 k = 0; k < g[s].n; ++k) if (an.gene_start[g[s].v[k]] <= bp && an.gene_end[g[s].v[k]] >= bp) pick = g[s].v[k];
						g[s].clear(); g[s].insert(pick); changed[s] = true;
					}
				}
				if (na == 3 && g[0].n && g[1].n && g[0].v[0] != g[1].v[0] && (an.gene_flags[g[0].v[0]] & GF_DUMMY) && (an.gene_flags[g[1].v[0]] & GF_DUMMY)) {
					const i32 bp = f.fwd(a[1]) ? f.start[a[1]] : f.end[a[1]];
					u32 pick = g[0].v[0];
					for (u32 k = 0; k < g[0].n; ++k) if (an.gene_start[g[0].v[k]] <= bp && an.gene_end[g[0].v[k]] >= bp) pick = g[0].v[k];
					for (u32 k = 0; k < g[1].n; ++k) if (an.gene_start[g[1].v[k]] <= bp && an.gene_end[g[1].v[k]] >= bp) pick = g[1].v[k];
					g[0].clear(); g[0].insert(pick); g[1].clear(); g[1].insert(pick); changed[0] = changed[1] = true;
				}
				for (u32 s = 0; s < na; ++s) if (changed[s]) { check_overflow(g[s]); refs.set(a[s], t, g[s].v, g[s].n); }
			}
		});
	}

	laps.lap("pass 2 (dummy genes)");
	// final CSR gene columns
	// offsets: per-slice sums on the threads, slice bases serially, then the prefix inside every slice
	const size_t A = 3 * (size_t) n;
	column<u64> at(A + 1);
	std::vector<u64> slice_sum(T + 1, 0);
	parallel_ranges(T, A, [&](int t, size_t lo, size_t hi) { u64 s = 0; for (size_t a = lo; a < hi; ++a) s += refs.cnt[a]; slice_sum[t + 1] = s; });
	for (int t = 0; t < T; ++t) slice_sum[t + 1] += slice_sum[t];
	parallel_ranges(T, A, [&](int t, size_t lo, size_t hi) { u64 s = slice_sum[t]; for (size_t a = lo; a < hi; ++a) { at[a] = s; s += refs.cnt[a]; } });
	at[A] = slice_sum[T];
	if (at[A] > 0xFFFFFFFFull) throw std::runtime_error("gene pool exceeds 2^32 entries");
	ft.genes.resize(at[A] + 1); ft.genes[at[A]] = 0;
	parallel_ranges(T, 3 * (size_t) n, [&](int, size_t lo, size_t hi) {
		for (size_t a = lo; a < hi; ++a) { ft.genes_off[a] = (u32) at[a]; ft.genes_cnt[a] = refs.cnt[a]; if (refs.cnt[a]) memcpy(&ft.genes[at[a]], refs.get(a), 4ull * refs.cnt[a]); }
	});
	laps.lap("gene columns");
}

// ------------------------------------------------------------------------------------------- fragment length
// read_stats.cpp:11-92. `early` = labels after the contig filters (what the reference's loop sees at arriba.cpp:357).
bool estimate_fragment_length(pipeline& p, const u8* early, float& gap_mean, float& gap_stddev, float& read_length_mean) {
	const annot_view an = p.ref.host_view(); const frag_view f = p.frags.view();
	std::list<int> gaps; unsigned int gap_count = 0, read_length_count = 0;
	read_length_mean = 0;
	for (u32 i = 0; i < f.n; ++i) {
		const u32 m = f.idx(i, MATE1), s = f.idx(i, MATE2);
		read_length_mean += ((size_t) f.seq_len[m] + (size_t) f.seq_len[s]) / 2; // integer division, then float accumulation
		++read_length_count;
		if (early[i] != F_none || (p.frags.fflags[i] & FF_SINGLE_END)) continue;
		if (f.n_aln[i] != 3) continue;
		u32 fm = m, rm = s;
		if (!f.fwd(fm)) { u32 t = fm; fm = rm; rm = t; }
		int distance = spliced_distance(an, f.contig[fm], f.end[fm], f.start[rm], f.genes[f.genes_off[fm]]);
		if (f.end[fm] > f.start[rm]) distance *= -1;
		if (distance < -(int) f.seq_len[fm]) distance = -(int) f.seq_len[fm];
		if (distance < -(int) f.seq_len[rm]) distance = -(int) f.seq_len[rm];
		gaps.push_back(distance);
		if (++gap_count > 100000) break;
	}
	if (gap_count < 10000) { std::cerr << "WARNING: not enough chimeric reads to estimate mate gap distribution, using default values" << std::endl; return false; }
	read_length_mean = read_length_mean / read_length_count;
	bool no_more_outliers = false;
	for (;;) {
		gap_mean = 0;
		for (std::list<int>::iterator g = gaps.begin(); g != gaps.end(); ++g) gap_mean += *g;
		gap_mean /= gap_count;
		gap_stddev = 0;
		for (std::list<int>::iterator g = gaps.begin(); g != gaps.end(); ++g) gap_stddev += (*g - gap_mean) * (*g - gap_mean);
		gap_stddev = sqrt(1.0 / (gap_count - 1) * gap_stddev);
		unsigned int within = 0;
		for (std::list<int>::iterator g = gaps.begin(); g != gaps.end(); ++g) if (*g > gap_mean - gap_stddev || *g < gap_mean + gap_stddev) ++within; // always true (read_stats.cpp:73)
		if (1.0 * within / gap_count < 0.683 || no_more_outliers) break;
		no_more_outliers = true;
		for (std::list<int>::iterator g = gaps.begin(); g != gaps.end();) {
			if (*g < gap_mean - 3 * gap_stddev || *g > gap_mean + 3 * gap_stddev) { g = gaps.erase(g); --gap_count; no_more_outliers = false; } else ++g;
		}
	}
	return true;
}

}} // namespace
