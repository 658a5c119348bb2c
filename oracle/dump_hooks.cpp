// oracle/dump_hooks.cpp -- TEST INFRASTRUCTURE (linked only into oracle/_ref/arriba).
//
// Observation hooks around the reference's own stage functions. The reference `main`
// (arriba.cpp:79-631) is compiled unmodified; GNU ld's --wrap interposes these wrappers on the calls
// main makes (arriba.cpp:130,143,153,329-413,422-589), each of which forwards to the real function
// and, when ARB_DUMP_DIR is set, writes the state of the containers (chimeric_alignments_t common.hpp:220,
// fusions_t common.hpp:286) to <dir>/<seq>_<stage>.bin. Without ARB_DUMP_DIR the binary behaves exactly
// like the reference CLI. The parity tests diff the product's per-stage state against these dumps.
//
// File format ("ARBD1"): repeated { u32 name_len, name, char dtype[2], u64 count, raw little-endian data }.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <unordered_map>
#include <vector>
#include "common.hpp"
#include "annotation.hpp"
#include "assembly.hpp"
#include "read_stats.hpp"
#include "filter_mismappers.hpp"
#include "annotate_tags.hpp"
#include "annotate_protein_domains.hpp"
#include "hook_syms.h"

using namespace std;

namespace {

const char* g_dir = getenv("ARB_DUMP_DIR");
int g_level = getenv("ARB_DUMP_LEVEL") ? atoi(getenv("ARB_DUMP_LEVEL")) : 2;
int g_seq = 0;
chimeric_alignments_t* g_frags = NULL;
unordered_map<const mates_t*, uint32_t> g_frag_index;
const exon_annotation_index_t* g_exon_index = NULL;

struct dump_t {
	FILE* f;
	explicit dump_t(const string& stage) {
		char path[4096];
		snprintf(path, sizeof(path), "%s/%02d_%s.bin", g_dir, g_seq++, stage.c_str());
		f = fopen(path, "wb");
		if (!f) { perror(path); exit(1); }
		fwrite("ARBD1\n", 1, 6, f);
	}
	~dump_t() { fclose(f); }
	void raw(const string& name, const char* dtype, const void* p, uint64_t count, size_t elem) {
		uint32_t nl = name.size();
		fwrite(&nl, 4, 1, f); fwrite(name.data(), 1, nl, f); fwrite(dtype, 1, 2, f); fwrite(&count, 8, 1, f);
		if (count) fwrite(p, elem, count, f);
	}
	void u1(const string& n, const vector<uint8_t>& v) { raw(n, "u1", v.data(), v.size(), 1); }
	void u2(const string& n, const vector<uint16_t>& v) { raw(n, "u2", v.data(), v.size(), 2); }
	void u4(const string& n, const vector<uint32_t>& v) { raw(n, "u4", v.data(), v.size(), 4); }
	void i4(const string& n, const vector<int32_t>& v) { raw(n, "i4", v.data(), v.size(), 4); }
	void u8(const string& n, const vector<uint64_t>& v) { raw(n, "u8", v.data(), v.size(), 8); }
	void f4(const string& n, const vector<float>& v) { raw(n, "f4", v.data(), v.size(), 4); }
	void f8(const string& n, const vector<double>& v) { raw(n, "f8", v.data(), v.size(), 8); }
	void str(const string& n, const string& s) { raw(n, "u1", s.data(), s.size(), 1); }
};

inline bool on() { return g_dir != NULL; }

void index_fragments(chimeric_alignments_t& ca) {
	g_frags = &ca;
	g_frag_index.clear();
	uint32_t i = 0;
	for (chimeric_alignments_t::iterator it = ca.begin(); it != ca.end(); ++it)
		g_frag_index[&it->second] = i++;
}

void dump_filter_bytes(dump_t& d, const chimeric_alignments_t& ca) {
	vector<uint8_t> filt; filt.reserve(ca.size());
	for (chimeric_alignments_t::const_iterator it = ca.begin(); it != ca.end(); ++it) filt.push_back(it->second.filter);
	d.u1("frag_filter", filt);
}

void dump_fragments(dump_t& d, const chimeric_alignments_t& ca) {
	vector<uint32_t> name_off(1, 0), aln_off(1, 0), cigar_off(1, 0), seq_off(1, 0), genes_off(1, 0), cigar, genes;
	string names, seqs;
	vector<uint8_t> n_aln, single_end, multimapper, duplicate, filt;
	vector<uint8_t> supp, first, exonic, strand, pstrand, pamb;
	vector<uint16_t> contig; vector<int32_t> start, end;
	for (chimeric_alignments_t::const_iterator it = ca.begin(); it != ca.end(); ++it) {
		names += it->first; name_off.push_back(names.size());
		const mates_t& m = it->second;
		n_aln.push_back(m.size()); single_end.push_back(m.single_end); multimapper.push_back(m.multimapper);
		duplicate.push_back(m.duplicate); filt.push_back(m.filter);
		for (mates_t::const_iterator a = m.begin(); a != m.end(); ++a) {
			supp.push_back(a->supplementary); first.push_back(a->first_in_pair); exonic.push_back(a->exonic);
			strand.push_back(a->strand); pstrand.push_back(a->predicted_strand); pamb.push_back(a->predicted_strand_ambiguous);
			contig.push_back(a->contig); start.push_back(a->start); end.push_back(a->end);
			for (size_t k = 0; k < a->cigar.size(); ++k) cigar.push_back(a->cigar[k]);
			cigar_off.push_back(cigar.size());
			seqs += a->sequence; seq_off.push_back(seqs.size());
			for (gene_set_t::const_iterator g = a->genes.begin(); g != a->genes.end(); ++g) genes.push_back((**g).id);
			genes_off.push_back(genes.size());
		}
		aln_off.push_back(supp.size());
	}
	d.str("names", names); d.u4("name_off", name_off);
	d.u1("n_aln", n_aln); d.u1("single_end", single_end); d.u1("multimapper", multimapper); d.u1("duplicate", duplicate); d.u1("frag_filter", filt);
	d.u4("aln_off", aln_off);
	d.u1("supplementary", supp); d.u1("first_in_pair", first); d.u1("exonic", exonic); d.u1("strand", strand);
	d.u1("predicted_strand", pstrand); d.u1("predicted_strand_ambiguous", pamb);
	d.u2("contig", contig); d.i4("start", start); d.i4("end", end);
	d.u4("cigar_off", cigar_off); d.u4("cigar", cigar);
	d.u4("seq_off", seq_off); d.str("seq", seqs);
	d.u4("genes_off", genes_off); d.u4("genes", genes);
}

template <class INDEX> void collect_genes(const INDEX& index, vector<gene_t>& out) {
	for (size_t c = 0; c < index.size(); ++c)
		for (typename INDEX::value_type::const_iterator r = index[c].begin(); r != index[c].end(); ++r)
			for (gene_set_t::const_iterator g = r->second.begin(); g != r->second.end(); ++g)
				out.push_back(*g);
	sort(out.begin(), out.end());
	out.erase(unique(out.begin(), out.end()), out.end());
}

void dump_genes(dump_t& d, const gene_annotation_index_t& index) {
	vector<gene_t> genes; // sorted by pointer value
	collect_genes(index, genes);
	vector<uint32_t> id, ptr_rank, name_off(1, 0), gid_off(1, 0); vector<uint16_t> contig; vector<int32_t> start, end, exonic_length;
	vector<uint8_t> strand, dummy, coding; string names, gids;
	vector<pair<unsigned int, size_t> > by_id;
	for (size_t i = 0; i < genes.size(); ++i) by_id.push_back(make_pair(genes[i]->id, i));
	sort(by_id.begin(), by_id.end());
	for (size_t k = 0; k < by_id.size(); ++k) {
		const gene_annotation_record_t& g = *genes[by_id[k].second];
		id.push_back(g.id); ptr_rank.push_back(by_id[k].second); contig.push_back(g.contig); start.push_back(g.start); end.push_back(g.end);
		exonic_length.push_back(g.exonic_length); strand.push_back(g.strand); dummy.push_back(g.is_dummy); coding.push_back(g.is_protein_coding);
		names += g.name; name_off.push_back(names.size()); gids += g.gene_id; gid_off.push_back(gids.size());
	}
	d.u4("gene_id", id); d.u4("gene_ptr_rank", ptr_rank); d.u2("gene_contig", contig); d.i4("gene_start", start); d.i4("gene_end", end);
	d.i4("gene_exonic_length", exonic_length); d.u1("gene_strand", strand); d.u1("gene_is_dummy", dummy); d.u1("gene_is_protein_coding", coding);
	d.str("gene_names", names); d.u4("gene_name_off", name_off); d.str("gene_gtf_ids", gids); d.u4("gene_gtf_id_off", gid_off);
}

void dump_exons(dump_t& d, const exon_annotation_index_t& index) {
	// the disjoint-region index exactly as the reference builds it (annotation.t.hpp:25-45) plus the exon records
	vector<exon_t> exons;
	vector<uint32_t> region_contig, region_off(1, 0), region_exons; vector<int32_t> region_end;
	for (size_t c = 0; c < index.size(); ++c)
		for (exon_contig_annotation_index_t::const_iterator r = index[c].begin(); r != index[c].end(); ++r)
			for (exon_set_t::const_iterator e = r->second.begin(); e != r->second.end(); ++e) exons.push_back(*e);
	sort(exons.begin(), exons.end());
	exons.erase(unique(exons.begin(), exons.end()), exons.end());
	unordered_map<exon_t, uint32_t> rank;
	for (size_t i = 0; i < exons.size(); ++i) rank[exons[i]] = i;
	for (size_t c = 0; c < index.size(); ++c)
		for (exon_contig_annotation_index_t::const_iterator r = index[c].begin(); r != index[c].end(); ++r) {
			region_contig.push_back(c); region_end.push_back(r->first);
			for (exon_set_t::const_iterator e = r->second.begin(); e != r->second.end(); ++e) region_exons.push_back(rank[*e]);
			region_off.push_back(region_exons.size());
		}
	vector<uint32_t> gene, transcript; vector<int32_t> start, end, cds_start, cds_end, prev, next;
	for (size_t i = 0; i < exons.size(); ++i) {
		const exon_annotation_record_t& e = *exons[i];
		gene.push_back(e.gene->id); transcript.push_back(e.transcript->id); start.push_back(e.start); end.push_back(e.end);
		cds_start.push_back(e.coding_region_start); cds_end.push_back(e.coding_region_end);
		prev.push_back(e.previous_exon ? (int32_t) rank[e.previous_exon] : -1); next.push_back(e.next_exon ? (int32_t) rank[e.next_exon] : -1);
	}
	d.u4("exon_gene", gene); d.u4("exon_transcript", transcript); d.i4("exon_start", start); d.i4("exon_end", end);
	d.i4("exon_cds_start", cds_start); d.i4("exon_cds_end", cds_end); d.i4("exon_prev", prev); d.i4("exon_next", next);
	d.u4("region_contig", region_contig); d.i4("region_end", region_end); d.u4("region_off", region_off); d.u4("region_exons", region_exons);
}

void dump_fusions(dump_t& d, const fusions_t& fusions, bool with_lists) {
	size_t n = fusions.size();
	vector<uint32_t> g1, g2, sr1, sr2, dm, l1_off(1, 0), l2_off(1, 0), ld_off(1, 0), l1, l2, ld;
	vector<uint16_t> c1, c2; vector<int32_t> b1, b2, a1, a2, cg1, cg2; vector<float> ev;
	vector<uint8_t> d1, d2, filt, ex1, ex2, sp1, sp2, ps1, ps2, psa, ts, tsa, conf;
	g1.reserve(n);
	for (fusions_t::const_iterator it = fusions.begin(); it != fusions.end(); ++it) {
		const fusion_t& f = it->second;
		g1.push_back(f.gene1->id); g2.push_back(f.gene2->id); c1.push_back(f.contig1); c2.push_back(f.contig2);
		b1.push_back(f.breakpoint1); b2.push_back(f.breakpoint2); d1.push_back(f.direction1); d2.push_back(f.direction2);
		sr1.push_back(f.split_reads1); sr2.push_back(f.split_reads2); dm.push_back(f.discordant_mates); filt.push_back(f.filter);
		ex1.push_back(f.exonic1); ex2.push_back(f.exonic2); sp1.push_back(f.spliced1); sp2.push_back(f.spliced2);
		ps1.push_back(f.predicted_strand1); ps2.push_back(f.predicted_strand2); psa.push_back(f.predicted_strands_ambiguous);
		ts.push_back(f.transcript_start); tsa.push_back(f.transcript_start_ambiguous); conf.push_back(f.confidence);
		ev.push_back(f.evalue); a1.push_back(f.anchor_start1); a2.push_back(f.anchor_start2);
		cg1.push_back(f.closest_genomic_breakpoint1); cg2.push_back(f.closest_genomic_breakpoint2);
		if (with_lists) {
			for (size_t k = 0; k < f.split_read1_list.size(); ++k) l1.push_back(g_frag_index.at(&f.split_read1_list[k]->second));
			for (size_t k = 0; k < f.split_read2_list.size(); ++k) l2.push_back(g_frag_index.at(&f.split_read2_list[k]->second));
			for (size_t k = 0; k < f.discordant_mate_list.size(); ++k) ld.push_back(g_frag_index.at(&f.discordant_mate_list[k]->second));
		}
		l1_off.push_back(with_lists ? l1.size() : f.split_read1_list.size() + l1_off.back());
		l2_off.push_back(with_lists ? l2.size() : f.split_read2_list.size() + l2_off.back());
		ld_off.push_back(with_lists ? ld.size() : f.discordant_mate_list.size() + ld_off.back());
	}
	d.u4("gene1", g1); d.u4("gene2", g2); d.u2("contig1", c1); d.u2("contig2", c2); d.i4("breakpoint1", b1); d.i4("breakpoint2", b2);
	d.u1("direction1", d1); d.u1("direction2", d2); d.u4("split_reads1", sr1); d.u4("split_reads2", sr2); d.u4("discordant_mates", dm);
	d.u1("filter", filt); d.u1("exonic1", ex1); d.u1("exonic2", ex2); d.u1("spliced1", sp1); d.u1("spliced2", sp2);
	d.u1("predicted_strand1", ps1); d.u1("predicted_strand2", ps2); d.u1("predicted_strands_ambiguous", psa);
	d.u1("transcript_start", ts); d.u1("transcript_start_ambiguous", tsa); d.u1("confidence", conf);
	d.f4("evalue", ev); d.i4("anchor_start1", a1); d.i4("anchor_start2", a2);
	d.i4("closest_genomic_breakpoint1", cg1); d.i4("closest_genomic_breakpoint2", cg2);
	d.u4("list1_off", l1_off); d.u4("list2_off", l2_off); d.u4("listd_off", ld_off);
	if (with_lists) { d.u4("list1", l1); d.u4("list2", l2); d.u4("listd", ld); }
}

// slot order of discordant mates may be canonicalised by find_fusions (fusions.cpp:416-421)
void dump_slot0(dump_t& d, const chimeric_alignments_t& ca) {
	vector<int32_t> s0; vector<uint16_t> c0;
	for (chimeric_alignments_t::const_iterator it = ca.begin(); it != ca.end(); ++it) { s0.push_back(it->second[MATE1].start); c0.push_back(it->second[MATE1].contig); }
	d.i4("slot0_start", s0); d.u2("slot0_contig", c0);
}

void after_read_filter(const char* stage, chimeric_alignments_t& ca, unsigned int remaining) {
	if (!on()) return;
	dump_t d(stage);
	d.u4("remaining", vector<uint32_t>(1, remaining));
	dump_filter_bytes(d, ca);
}

void after_event_filter(const char* stage, fusions_t& fusions, unsigned int remaining, bool with_lists = false, bool with_frag_filters = false) {
	if (!on()) return;
	dump_t d(stage);
	d.u4("remaining", vector<uint32_t>(1, remaining));
	dump_fusions(d, fusions, with_lists && g_level >= 2);
	if (with_frag_filters && g_frags) dump_filter_bytes(d, *g_frags);
}

} // namespace

#define HOOK(ret, name, params) \
	ret real_##name params asm("__real_" MANGLED_##name); \
	ret wrap_##name params asm("__wrap_" MANGLED_##name); \
	ret wrap_##name params

// ------------------------------------------------------------------ ingest
HOOK(unsigned int, read_chimeric_alignments, (const string& bam_file_path, const assembly_t& assembly, const string& assembly_file_path, chimeric_alignments_t& chimeric_alignments, unsigned long int& mapped_reads, vector<unsigned long int>& mapped_viral_reads_by_contig, coverage_t& coverage, contigs_t& contigs, vector<string>& original_contig_names, const string& interesting_contigs, const string& viral_contigs, const gene_annotation_index_t& gene_annotation_index, const bool separate_chimeric_bam_file, const bool is_rna_bam_file, const bool external_duplicate_marking, const unsigned int max_itd_length, const int threads)) {
	unsigned int r = real_read_chimeric_alignments(bam_file_path, assembly, assembly_file_path, chimeric_alignments, mapped_reads, mapped_viral_reads_by_contig, coverage, contigs, original_contig_names, interesting_contigs, viral_contigs, gene_annotation_index, separate_chimeric_bam_file, is_rna_bam_file, external_duplicate_marking, max_itd_length, threads);
	if (on()) {
		dump_t d("read_chimeric_alignments");
		d.u8("mapped_reads", vector<uint64_t>(1, mapped_reads));
		d.u4("total", vector<uint32_t>(1, r));
		string names; vector<uint32_t> off(1, 0);
		for (size_t c = 0; c < original_contig_names.size(); ++c) { names += original_contig_names[c]; off.push_back(names.size()); }
		d.str("contig_names", names); d.u4("contig_name_off", off);
		if (g_level >= 2) {
			vector<uint32_t> cov_off(1, 0); vector<uint16_t> cov; vector<uint8_t> starts, ends;
			for (size_t c = 0; c < coverage.coverage.size(); ++c) {
				cov.insert(cov.end(), coverage.coverage[c].begin(), coverage.coverage[c].end());
				for (size_t i = 0; i < coverage.fragment_starts[c].size(); ++i) { starts.push_back(coverage.fragment_starts[c][i]); ends.push_back(coverage.fragment_ends[c][i]); }
				cov_off.push_back(cov.size());
			}
			d.u4("coverage_off", cov_off); d.u2("coverage", cov); d.u1("fragment_starts", starts); d.u1("fragment_ends", ends);
		}
	}
	return r;
}

HOOK(unsigned int, mark_multimappers, (chimeric_alignments_t& chimeric_alignments)) {
	unsigned int r = real_mark_multimappers(chimeric_alignments);
	if (on()) {
		index_fragments(chimeric_alignments);
		dump_t d("ingest");
		d.u4("marked", vector<uint32_t>(1, r));
		if (g_level >= 2) dump_fragments(d, chimeric_alignments);
	}
	return r;
}

HOOK(strandedness_t, detect_strandedness, (const chimeric_alignments_t& chimeric_alignments, const gene_annotation_index_t& gene_annotation_index, const exon_annotation_index_t& exon_annotation_index)) {
	strandedness_t r = real_detect_strandedness(chimeric_alignments, gene_annotation_index, exon_annotation_index);
	if (on()) { dump_t d("strandedness"); d.u1("strandedness", vector<uint8_t>(1, (uint8_t) r)); }
	return r;
}

// ------------------------------------------------------------------ read-level cascade (arriba.cpp:327-409)
HOOK(unsigned int, filter_duplicates, (chimeric_alignments_t& chimeric_alignments, const bool external_duplicate_marking)) {
	if (on()) { // state after annotation (arriba.cpp:165-325), before the first filter
		if (!g_frags) index_fragments(chimeric_alignments);
		dump_t d("annotated");
		if (g_level >= 2) dump_fragments(d, chimeric_alignments);
	}
	unsigned int r = real_filter_duplicates(chimeric_alignments, external_duplicate_marking);
	after_read_filter("rf_duplicates", chimeric_alignments, r);
	return r;
}
HOOK(unsigned int, filter_uninteresting_contigs, (chimeric_alignments_t& ca, const vector<bool>& interesting_contigs)) {
	unsigned int r = real_filter_uninteresting_contigs(ca, interesting_contigs); after_read_filter("rf_uninteresting_contigs", ca, r); return r;
}
HOOK(unsigned int, filter_viral_contigs, (chimeric_alignments_t& ca, const vector<bool>& viral_contigs)) {
	unsigned int r = real_filter_viral_contigs(ca, viral_contigs); after_read_filter("rf_viral_contigs", ca, r); return r;
}
HOOK(unsigned int, filter_top_expressed_viral_contigs, (chimeric_alignments_t& ca, unsigned int top_count, const vector<bool>& viral_contigs, const vector<bool>& interesting_contigs, const vector<unsigned long int>& mapped_viral_reads_by_contig, const assembly_t& assembly)) {
	unsigned int r = real_filter_top_expressed_viral_contigs(ca, top_count, viral_contigs, interesting_contigs, mapped_viral_reads_by_contig, assembly); after_read_filter("rf_top_expressed_viral_contigs", ca, r); return r;
}
HOOK(unsigned int, filter_low_coverage_viral_contigs, (chimeric_alignments_t& ca, const coverage_t& coverage, const vector<bool>& viral_contigs, const float min_covered_fraction, const float min_covered_bases)) {
	unsigned int r = real_filter_low_coverage_viral_contigs(ca, coverage, viral_contigs, min_covered_fraction, min_covered_bases); after_read_filter("rf_low_coverage_viral_contigs", ca, r); return r;
}
HOOK(bool, estimate_fragment_length, (const chimeric_alignments_t& ca, float& mate_gap_mean, float& mate_gap_stddev, float& read_length_mean, const gene_annotation_index_t& gene_annotation_index, const exon_annotation_index_t& exon_annotation_index)) {
	bool r = real_estimate_fragment_length(ca, mate_gap_mean, mate_gap_stddev, read_length_mean, gene_annotation_index, exon_annotation_index);
	g_exon_index = &exon_annotation_index;
	if (on()) {
		dump_t d("fragment_length");
		d.u1("ok", vector<uint8_t>(1, r));
		vector<float> v; v.push_back(r ? mate_gap_mean : 0); v.push_back(r ? mate_gap_stddev : 0); v.push_back(read_length_mean);
		d.f4("gap_mean_stddev_readlen", v);
		dump_genes(d, gene_annotation_index);
		if (g_level >= 2) dump_exons(d, exon_annotation_index);
	}
	return r;
}
HOOK(unsigned int, filter_proximal_read_through, (chimeric_alignments_t& ca, const int min_distance)) {
	unsigned int r = real_filter_proximal_read_through(ca, min_distance); after_read_filter("rf_read_through", ca, r); return r;
}
HOOK(unsigned int, filter_inconsistently_clipped_mates, (chimeric_alignments_t& ca)) {
	unsigned int r = real_filter_inconsistently_clipped_mates(ca); after_read_filter("rf_inconsistently_clipped", ca, r); return r;
}
HOOK(unsigned int, filter_homopolymer, (chimeric_alignments_t& ca, const unsigned int homopolymer_length, const exon_annotation_index_t& exon_annotation_index)) {
	unsigned int r = real_filter_homopolymer(ca, homopolymer_length, exon_annotation_index); after_read_filter("rf_homopolymer", ca, r); return r;
}
HOOK(unsigned int, filter_small_insert_size, (chimeric_alignments_t& ca, const unsigned int max_overhang)) {
	unsigned int r = real_filter_small_insert_size(ca, max_overhang); after_read_filter("rf_small_insert_size", ca, r); return r;
}
HOOK(unsigned int, filter_long_gap, (chimeric_alignments_t& ca)) {
	unsigned int r = real_filter_long_gap(ca); after_read_filter("rf_long_gap", ca, r); return r;
}
HOOK(unsigned int, filter_same_gene, (chimeric_alignments_t& ca, exon_annotation_index_t& exon_annotation_index)) {
	unsigned int r = real_filter_same_gene(ca, exon_annotation_index); after_read_filter("rf_same_gene", ca, r); return r;
}
HOOK(unsigned int, filter_hairpin, (chimeric_alignments_t& ca, exon_annotation_index_t& exon_annotation_index, const int max_mate_gap)) {
	unsigned int r = real_filter_hairpin(ca, exon_annotation_index, max_mate_gap);
	if (on()) { dump_t d("rf_hairpin"); d.u4("remaining", vector<uint32_t>(1, r)); d.i4("max_mate_gap", vector<int32_t>(1, max_mate_gap)); dump_filter_bytes(d, ca); }
	return r;
}
HOOK(unsigned int, filter_mismatches, (chimeric_alignments_t& ca, const assembly_t& assembly, const vector<bool>& interesting_contigs, const vector<bool>& viral_contigs, const float mismatch_probability, const float pvalue_cutoff)) {
	unsigned int r = real_filter_mismatches(ca, assembly, interesting_contigs, viral_contigs, mismatch_probability, pvalue_cutoff); after_read_filter("rf_mismatches", ca, r); return r;
}
HOOK(unsigned int, filter_low_entropy, (chimeric_alignments_t& ca, const unsigned int kmer_length, const float kmer_content, const unsigned int max_itd_length)) {
	unsigned int r = real_filter_low_entropy(ca, kmer_length, kmer_content, max_itd_length); after_read_filter("rf_low_entropy", ca, r); return r;
}

// ------------------------------------------------------------------ candidates (arriba.cpp:411-589)
HOOK(unsigned int, find_fusions, (chimeric_alignments_t& ca, fusions_t& fusions, exon_annotation_index_t& exon_annotation_index, const int max_mate_gap, const unsigned int subsampling_threshold)) {
	unsigned int r = real_find_fusions(ca, fusions, exon_annotation_index, max_mate_gap, subsampling_threshold);
	if (on()) {
		dump_t d("find_fusions");
		d.u4("remaining", vector<uint32_t>(1, r));
		d.i4("max_mate_gap", vector<int32_t>(1, max_mate_gap));
		dump_fusions(d, fusions, g_level >= 2);
		dump_slot0(d, ca);
	}
	return r;
}
HOOK(unsigned int, merge_adjacent_fusions, (fusions_t& fusions, const int max_distance, const unsigned int max_itd_length)) {
	unsigned int r = real_merge_adjacent_fusions(fusions, max_distance, max_itd_length); after_event_filter("ev_merge_adjacent", fusions, r, true); return r;
}
HOOK(unsigned int, filter_multimappers, (chimeric_alignments_t& ca, fusions_t& fusions, const exon_annotation_index_t& exon_annotation_index, const assembly_t& assembly)) {
	unsigned int r = real_filter_multimappers(ca, fusions, exon_annotation_index, assembly); after_event_filter("ev_multimappers", fusions, r, false, true); return r;
}
HOOK(void, estimate_expected_fusions, (fusions_t& fusions, const unsigned long int mapped_reads, const exon_annotation_index_t& exon_annotation_index)) {
	real_estimate_expected_fusions(fusions, mapped_reads, exon_annotation_index); after_event_filter("ev_evalue", fusions, 0);
}
HOOK(unsigned int, filter_non_coding_neighbors, (fusions_t& fusions)) {
	unsigned int r = real_filter_non_coding_neighbors(fusions); after_event_filter("ev_non_coding_neighbors", fusions, r); return r;
}
HOOK(unsigned int, filter_intragenic_both_exonic, (fusions_t& fusions, const exon_annotation_index_t& exon_annotation_index, const float exonic_fraction)) {
	unsigned int r = real_filter_intragenic_both_exonic(fusions, exon_annotation_index, exonic_fraction); after_event_filter("ev_intragenic_exonic", fusions, r); return r;
}
HOOK(unsigned int, filter_min_support, (fusions_t& fusions, const int min_support)) {
	unsigned int r = real_filter_min_support(fusions, min_support); after_event_filter("ev_min_support", fusions, r); return r;
}
HOOK(unsigned int, filter_relative_support, (fusions_t& fusions, const float evalue_cutoff)) {
	unsigned int r = real_filter_relative_support(fusions, evalue_cutoff); after_event_filter("ev_relative_support", fusions, r); return r;
}
HOOK(unsigned int, recover_internal_tandem_duplication, (fusions_t& fusions, const chimeric_alignments_t& ca, const coverage_t& coverage, const exon_annotation_index_t& exon_annotation_index, const unsigned int max_itd_length, const unsigned int min_supporting_reads, const float min_fraction_of_coverage, const unsigned int subsampling_threshold)) {
	unsigned int r = real_recover_internal_tandem_duplication(fusions, ca, coverage, exon_annotation_index, max_itd_length, min_supporting_reads, min_fraction_of_coverage, subsampling_threshold);
	after_event_filter("ev_internal_tandem_duplication", fusions, r, true, true); return r;
}
HOOK(unsigned int, filter_both_intronic, (fusions_t& fusions, const vector<bool>& viral_contigs)) {
	unsigned int r = real_filter_both_intronic(fusions, viral_contigs); after_event_filter("ev_intronic", fusions, r); return r;
}
HOOK(unsigned int, filter_in_vitro, (fusions_t& fusions, const chimeric_alignments_t& ca, const float high_expression_quantile, const gene_annotation_index_t& gene_annotation_index, const coverage_t& coverage)) {
	unsigned int r = real_filter_in_vitro(fusions, ca, high_expression_quantile, gene_annotation_index, coverage); after_event_filter("ev_in_vitro", fusions, r); return r;
}
HOOK(unsigned int, recover_both_spliced, (fusions_t& fusions, const chimeric_alignments_t& ca, const exon_annotation_index_t& exon_annotation_index, const coverage_t& coverage, const unsigned int max_fusions_to_recover, const float high_expression_quantile, const int max_exon_size, const unsigned int max_coverage)) {
	unsigned int r = real_recover_both_spliced(fusions, ca, exon_annotation_index, coverage, max_fusions_to_recover, high_expression_quantile, max_exon_size, max_coverage); after_event_filter("ev_spliced", fusions, r); return r;
}
HOOK(unsigned int, select_most_supported_breakpoints, (fusions_t& fusions)) {
	unsigned int r = real_select_most_supported_breakpoints(fusions); after_event_filter("ev_select_best", fusions, r); return r;
}
HOOK(unsigned int, filter_marginal_read_through, (fusions_t& fusions, const coverage_t& coverage)) {
	unsigned int r = real_filter_marginal_read_through(fusions, coverage); after_event_filter("ev_marginal_read_through", fusions, r); return r;
}
HOOK(unsigned int, recover_many_spliced, (fusions_t& fusions, const unsigned int min_spliced_events)) {
	unsigned int r = real_recover_many_spliced(fusions, min_spliced_events); after_event_filter("ev_many_spliced", fusions, r); return r;
}
HOOK(unsigned int, filter_short_anchor, (fusions_t& fusions, unsigned int min_length)) {
	unsigned int r = real_filter_short_anchor(fusions, min_length); after_event_filter("ev_short_anchor", fusions, r); return r;
}
HOOK(unsigned int, filter_end_to_end_fusions, (fusions_t& fusions, const exon_annotation_index_t& exon_annotation_index, const vector<bool>& viral_contigs)) {
	unsigned int r = real_filter_end_to_end_fusions(fusions, exon_annotation_index, viral_contigs); after_event_filter("ev_end_to_end", fusions, r); return r;
}
HOOK(unsigned int, filter_no_coverage, (fusions_t& fusions, const coverage_t& coverage, const exon_annotation_index_t& exon_annotation_index)) {
	unsigned int r = real_filter_no_coverage(fusions, coverage, exon_annotation_index); after_event_filter("ev_no_coverage", fusions, r); return r;
}
HOOK(void, make_kmer_index, (const fusions_t& fusions, const assembly_t& assembly, int padding, const char kmer_length, kmer_indices_t& kmer_indices)) {
	real_make_kmer_index(fusions, assembly, padding, kmer_length, kmer_indices);
	if (on()) {
		dump_t d("kmer_index");
		d.i4("padding", vector<int32_t>(1, padding));
		vector<uint64_t> n_kmers, n_pos, checksum;
		for (size_t c = 0; c < kmer_indices.size(); ++c) {
			uint64_t np = 0, cs = 0;
			for (kmer_index_t::const_iterator k = kmer_indices[c].begin(); k != kmer_indices[c].end(); ++k)
				for (size_t i = 0; i < k->second.size(); ++i) { ++np; cs += ((uint64_t) k->first * 1000003ULL + (uint64_t) k->second[i]) * 0x9E3779B97F4A7C15ULL; }
			n_kmers.push_back(kmer_indices[c].size()); n_pos.push_back(np); checksum.push_back(cs);
		}
		d.u8("kmers_per_contig", n_kmers); d.u8("positions_per_contig", n_pos); d.u8("checksum_per_contig", checksum);
	}
}
HOOK(unsigned int, filter_homologs, (fusions_t& fusions, const kmer_indices_t& kmer_indices, const char kmer_length, const assembly_t& assembly, const float max_identity_fraction)) {
	unsigned int r = real_filter_homologs(fusions, kmer_indices, kmer_length, assembly, max_identity_fraction); after_event_filter("ev_homologs", fusions, r); return r;
}
HOOK(unsigned int, filter_mismappers, (fusions_t& fusions, const kmer_indices_t& kmer_indices, const char kmer_length, const assembly_t& assembly, const exon_annotation_index_t& exon_annotation_index, const float max_mismapper_fraction, const int max_mate_gap)) {
	unsigned int r = real_filter_mismappers(fusions, kmer_indices, kmer_length, assembly, exon_annotation_index, max_mismapper_fraction, max_mate_gap);
	after_event_filter("ev_mismappers", fusions, r, false, true); return r;
}
HOOK(unsigned int, recover_isoforms, (fusions_t& fusions)) {
	unsigned int r = real_recover_isoforms(fusions); after_event_filter("ev_isoforms", fusions, r); return r;
}
HOOK(void, assign_confidence, (fusions_t& fusions, const coverage_t& coverage)) {
	real_assign_confidence(fusions, coverage); after_event_filter("ev_confidence", fusions, 0, true, true);
}
HOOK(void, write_fusions_to_file, (fusions_t& fusions, const string& output_file, const coverage_t& coverage, const assembly_t& assembly, gene_annotation_index_t& gene_annotation_index, exon_annotation_index_t& exon_annotation_index, vector<string> original_contig_names, const tags_t& tags, const protein_domain_annotation_index_t& protein_domain_annotation_index, const int max_mate_gap, const unsigned int max_itd_length, const bool print_extra_info, const bool fill_sequence_gaps, const bool write_discarded_fusions)) {
	real_write_fusions_to_file(fusions, output_file, coverage, assembly, gene_annotation_index, exon_annotation_index, original_contig_names, tags, protein_domain_annotation_index, max_mate_gap, max_itd_length, print_extra_info, fill_sequence_gaps, write_discarded_fusions);
}
