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

typedef idset<1024> gset;       // gene set of one query on the host (strandedness detection)
template <int CAP> static void check_overflow(const idset<CAP>& s) { if (s.overflow) throw std::runtime_error("more than " + std::to_string(CAP) + " overlapping genes or exons under one alignment are not supported"); }
static void check(arb_ctx* ctx, int rc, const char* what) { if (rc != 0) throw std::runtime_error(std::string(what) + ": " + arb_last_error(ctx)); }

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

// ------------------------------------------------------------------------------------------- annotation passes
// The passes run on the device over the resident columns (csrc/annotate_hd.h, csrc/annotate.cu); the host only owns the gene table: it appends the dummy
// genes the device found between the passes and rebuilds the gene index over them (arriba.cpp:207-260).
void annotate_fragments(pipeline& p) {
	refdata& ref = p.ref;
	fragment_table& ft = p.frags;
	stage_laps laps("annotate");
	p.send_annotation();
	laps.lap("gene table -> device");
	u32 n_dummy = 0;
	check(p.ctx, arb_annotate_pass1(p.ctx, ft.aflags.data(), p.strandedness, &n_dummy), "arb_annotate_pass1");
	laps.lap("pass 1 + dummy gene clusters (device)");
	if (n_dummy) {
		std::vector<u16> contig(n_dummy); std::vector<i32> start(n_dummy), end(n_dummy);
		check(p.ctx, arb_get_dummy_genes(p.ctx, contig.data(), start.data(), end.data()), "arb_get_dummy_genes");
		gene_rec d; d.forward = true; d.exonic_length = 10000; d.is_dummy = true; d.is_protein_coding = false;
		ref.genes.reserve(ref.genes.size() + n_dummy);
		for (u32 k = 0; k < n_dummy; ++k) { d.contig = contig[k]; d.start = start[k]; d.end = end[k]; ref.genes.push_back(d); }
	}
	ref.build_gene_index();
	ref.flatten();
	laps.lap("gene index rebuilt");
	p.send_annotation();
	u64 n_gene_ids = 0;
	check(p.ctx, arb_annotate_pass2(p.ctx, &n_gene_ids), "arb_annotate_pass2");
	laps.lap("pass 2 + gene columns (device)");
	ft.genes.resize(n_gene_ids + 1); ft.genes[n_gene_ids] = 0;
	check(p.ctx, arb_get_annotation_columns(p.ctx, ft.aflags.data(), ft.genes_off.data(), ft.genes_cnt.data(), ft.genes.data()), "arb_get_annotation_columns");
	laps.lap("annotation columns -> host");
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
