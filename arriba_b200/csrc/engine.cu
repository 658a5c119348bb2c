// engine.cu -- stage drivers: uploads, duplicate marking, the fused read-level cascade.
#include <cstdlib>
#include <algorithm>
#include "engine.h"
#include "mismatch_table.h"

namespace arb {

void default_params(arb_params& p) { // options.cpp:71-107
	p.filter_mask = 0;
	for (u32 f = 1; f < F_COUNT; ++f) p.filter_mask |= (u64) 1 << f;
	p.homopolymer_length = 6; p.min_read_through_distance = 10000; p.max_kmer_content = 0.6f; p.max_itd_length = 100;
	p.external_duplicate_marking = 0; p.mismatch_pvalue_cutoff = 0.01f; p.subsampling_threshold = 300; p.evalue_cutoff = 0.3f;
	p.max_mismapper_fraction = 0.8f; p.max_homolog_identity = 0.3f;
}

engine::engine(): device(0), table_n(0), table_k(0), has_contigs(false), has_annotation(false), filters_done(false), merge_log_n(0), kmer_index_contigs(0), kmer_indexed(0), has_splice_sites(false) {
	default_params(params);
	memset(&timings, 0, sizeof(timings));
	// tuning hooks of the re-alignment passes (mismap_hd.h): ARB_MISMAP_BUDGET (0 = thread-per-item only), ARB_MISMAP_LANES, ARB_MISMAP_SPAWN (0 = no task rounds), ARB_MISMAP_TASK_LANES
	mismap_budget = 4096; mismap_lanes = 1024; mismap_spawn_budget = 0; mismap_task_lanes = 32;
	if (const char* s = getenv("ARB_MISMAP_BUDGET")) mismap_budget = atoi(s);
	if (const char* s = getenv("ARB_MISMAP_LANES")) mismap_lanes = (u32) std::max(1, atoi(s));
	if (const char* s = getenv("ARB_MISMAP_SPAWN")) mismap_spawn_budget = atoi(s);
	homolog_lanes = 32; if (const char* s = getenv("ARB_HOMOLOG_LANES")) homolog_lanes = (u32) std::max(1, atoi(s)); // threads per gene pair in filter_homologs' identity test (1 = one thread per pair)
	mismap_table_slots = 4096; if (const char* s = getenv("ARB_MISMAP_TABLE")) mismap_table_slots = (u32) std::max(1, atoi(s)); // continuation registry: slots provisioned per cooperative item (the table is shared, 2^20..2^25 slots)
	mismap_min_blocks = 4; if (const char* s = getenv("ARB_MISMAP_OCC")) mismap_min_blocks = atoi(s); // resident 256-thread blocks per SM the re-alignment kernels are compiled for
	if (const char* s = getenv("ARB_MISMAP_TASK_LANES")) mismap_task_lanes = (u32) std::max(1, atoi(s));
	ex.scratch = &scratch;
#ifdef ARB_DEVICE_BUILD
	ARB_CUDA_CHECK(cudaStreamCreateWithFlags(&ex.stream, cudaStreamNonBlocking));
#endif
}

engine::~engine() { // arb_ctx_destroy has made the context's device current; buffers go back to that device's pool after the stream has drained
#ifdef ARB_DEVICE_BUILD
	if (ex.stream) { cudaStreamSynchronize(ex.stream); cudaStreamDestroy(ex.stream); ex.stream = 0; }
#endif
}

// reference bases as nt16 codes, eight per word: the mismatch rule compares eight read bases per XOR (read_filters.h)
struct pack_assembly_fn {
	const char* bases; u32 length; u32* packed; u32* exotic; // one contig; bases and packed start at the contig's (64-base aligned) offset
	ARB_HD void operator()(u32 w) const {
		u32 word = 0;
		for (u32 b = 0; b < 8 && w * 8 + b < length; ++b) {
			const u32 code = nt16_of_char(bases[w * 8 + b]);
			if (code > 15) { *exotic = 1; continue; }
			word |= code << (28 - 4 * b);
		}
		packed[w] = word;
	}
};

void engine::set_contigs(const arb_contigs& c) {
	annot.n_contigs = c.n_contigs;
	annot.h_contig_flags.assign(c.flags, c.flags + c.n_contigs);
	annot.h_contig_len.assign(c.length, c.length + c.n_contigs);
	std::vector<u64> off(c.n_contigs);
	u64 total = 0;
	for (u32 k = 0; k < c.n_contigs; ++k) {
		const bool loaded = c.sequence && c.sequence[k] && c.length[k] > 0;
		off[k] = loaded ? total : ~(u64) 0;
		if (!loaded) annot.h_contig_len[k] = 0;
		if (loaded) total += ((u64) c.length[k] + 63) & ~(u64) 63;
	}
	annot.assembly.ensure(total + 64);
	for (u32 k = 0; k < c.n_contigs; ++k) {
		if (off[k] == ~(u64) 0) continue;
#ifdef ARB_DEVICE_BUILD
		ARB_CUDA_CHECK(cudaMemcpyAsync(annot.assembly.ptr() + off[k], c.sequence[k], c.length[k], cudaMemcpyHostToDevice, ex.stream));
#else
		memcpy(annot.assembly.ptr() + off[k], c.sequence[k], c.length[k]);
#endif
	}
	for (u32 k = 0; k < c.n_contigs; ++k) if (off[k] == ~(u64) 0) off[k] = 0; // contig_len == 0 guards every access
	annot.contig_flags.upload(ex, annot.h_contig_flags.data(), c.n_contigs);
	annot.contig_len.upload(ex, annot.h_contig_len.data(), c.n_contigs);
	annot.contig_seq_off.upload(ex, off.data(), c.n_contigs);
	{
		const u64 n_words = total / 8 + 9; // windows read one word past the last base
		annot.assembly4.ensure(n_words); annot.assembly4.zero(ex, n_words);
		dbuf<u32> exotic(1); exotic.zero(ex, 1);
		for (u32 k = 0; k < c.n_contigs; ++k) {
			if (annot.h_contig_len[k] == 0) continue;
			pack_assembly_fn pf = {annot.assembly.ptr() + off[k], annot.h_contig_len[k], annot.assembly4.ptr() + off[k] / 8, exotic.ptr()};
			for_each(ex, (annot.h_contig_len[k] + 7) / 8, pf);
		}
		u32 flag = 0; exotic.download(ex, &flag, 1);
		annot.assembly4_ok = flag == 0;
	}
	ex.sync();
	has_contigs = true;
	table_n = 0; // mismatch table depends on the genome size
}

void engine::set_annotation(const arb_annotation& a) {
	if (has_contigs && a.n_contigs != annot.n_contigs) throw arb_error("annotation and contig table disagree on the number of contigs");
	annot.n_genes = a.n_genes; annot.n_exons = a.n_exons; annot.n_contigs = a.n_contigs;
	annot.gene_contig.upload(ex, a.gene_contig, a.n_genes); annot.gene_start.upload(ex, a.gene_start, a.n_genes); annot.gene_end.upload(ex, a.gene_end, a.n_genes);
	annot.gene_strand.upload(ex, a.gene_strand, a.n_genes); annot.gene_exonic_length.upload(ex, a.gene_exonic_length, a.n_genes); annot.gene_flags.upload(ex, a.gene_flags, a.n_genes);
	annot.exon_gene.upload(ex, a.exon_gene, a.n_exons); annot.exon_start.upload(ex, a.exon_start, a.n_exons); annot.exon_end.upload(ex, a.exon_end, a.n_exons);
	annot.exon_cds_start.upload(ex, a.exon_cds_start, a.n_exons); annot.exon_cds_end.upload(ex, a.exon_cds_end, a.n_exons);
	annot.exon_next_start.upload(ex, a.exon_next_start, a.n_exons); annot.exon_flags.upload(ex, a.exon_flags, a.n_exons);
	const u32 ner = a.exon_region_begin[a.n_contigs], ngr = a.gene_region_begin[a.n_contigs];
	annot.exon_region_begin.upload(ex, a.exon_region_begin, a.n_contigs + 1); annot.exon_region_end.upload(ex, a.exon_region_end, ner);
	annot.exon_region_off.upload(ex, a.exon_region_off, ner + 1); annot.exon_region_items.upload(ex, a.exon_region_items, a.exon_region_off[ner]);
	annot.gene_region_begin.upload(ex, a.gene_region_begin, a.n_contigs + 1); annot.gene_region_end.upload(ex, a.gene_region_end, ngr);
	annot.gene_region_off.upload(ex, a.gene_region_off, ngr + 1); annot.gene_region_items.upload(ex, a.gene_region_items, a.gene_region_off[ngr]);
	annot.h_gene_contig.assign(a.gene_contig, a.gene_contig + a.n_genes); annot.h_gene_start.assign(a.gene_start, a.gene_start + a.n_genes);
	annot.h_gene_end.assign(a.gene_end, a.gene_end + a.n_genes); annot.h_gene_strand.assign(a.gene_strand, a.gene_strand + a.n_genes);
	annot.h_gene_flags.assign(a.gene_flags, a.gene_flags + a.n_genes);
	ex.sync();
	has_annotation = true;
}

void engine::push_chunk(const arb_soa_chunk& c) {
	const u32 n = c.n_fragments;
	frags.n = n;
	stage_timer t_h2d(ex);
	frags.n_aln.upload(ex, c.n_aln, n); frags.fflags.upload(ex, c.fflags, n); frags.filter.upload(ex, c.filter, n);
	frags.early.ensure(n); frags.swapped.ensure(n); frags.swapped.zero(ex, n);
	frags.contig.upload(ex, c.contig, 3 * (size_t) n); frags.start.upload(ex, c.start, 3 * (size_t) n); frags.end.upload(ex, c.end, 3 * (size_t) n);
	frags.aflags.upload(ex, c.aflags, 3 * (size_t) n); frags.cigar_off.upload(ex, c.cigar_off, 3 * (size_t) n); frags.cigar_cnt.upload(ex, c.cigar_cnt, 3 * (size_t) n);
	frags.seq_off.upload(ex, c.seq_off, 2 * (size_t) n); frags.seq_len.upload(ex, c.seq_len, 2 * (size_t) n);
	frags.genes_off.upload(ex, c.genes_off, 3 * (size_t) n); frags.genes_cnt.upload(ex, c.genes_cnt, 3 * (size_t) n);
	frags.cigar.upload(ex, c.cigar, c.n_cigar); frags.seq.upload(ex, c.seq, c.n_seq_bytes); frags.genes.upload(ex, c.genes, c.n_genes);
	u32 max_len = 0;
	for (size_t k = 0; k < 2 * (size_t) n; ++k) if (c.seq_len[k] > max_len) max_len = c.seq_len[k];
	frags.max_seq_len = max_len;
	timings.h2d_ms = t_h2d.stop();
	timings.h2d_bytes = (u64) n * 3 + (u64) n * 3 * (2 + 4 + 4 + 1 + 4 + 2 + 4 + 2) + (u64) n * 2 * (4 + 2) + c.n_cigar * 4 + c.n_seq_bytes + c.n_genes * 4;
	// Column budget of SURVEY.md section 8(d): every column read once at its compact width -- 11 B per alignment {contig u16, start, end, flags u8}, CIGAR ops and
	// gene ids with a 4-byte offset per alignment, 6 B per fragment {rank, flags, label}; sequences at 3 bit/base, gathered reference bases at 2 bit/base.
	u64 n_alignments = 0, bases = 0;
	for (size_t k = 0; k < n; ++k) n_alignments += c.n_aln[k];
	for (size_t k = 0; k < 2 * (size_t) n; ++k) bases += c.seq_len[k];
	head_bytes = n_alignments * 11 + ((u64) c.n_cigar + n_alignments) * 4 + ((u64) c.n_genes + n_alignments) * 4 + (u64) n * 6;
	sequence_bytes = bases * 3 / 8 + bases * 2 / 8 + ((u64) c.n_cigar + n_alignments) * 4 + n_alignments * 11 + (u64) n * 2;
	timings.classify_algorithmic_bytes = head_bytes + sequence_bytes;
	ex.sync();
	filters_done = false;
	cands.n = 0;
}

unsigned long engine::genome_size() const { // filter_mismatches.cpp:105-108
	unsigned long g = 0;
	for (u32 k = 0; k < annot.n_contigs; ++k) if (annot.h_contig_flags[k] & CF_INTERESTING) g += annot.h_contig_len[k];
	return g;
}

read_filter_params engine::make_filter_params() {
	// decision table large enough for every (aligned bases, mismatches) the resident fragments can produce
	const u32 need_n = frags.max_seq_len + 2, need_k = frags.max_seq_len + 70;
	if (table_n < need_n || table_k < need_k) {
		table_n = need_n; table_k = need_k;
		std::vector<u8> t = build_mismatch_table(table_n, table_k, 0.01f /* arriba.cpp:403 */, genome_size(), params.mismatch_pvalue_cutoff);
		mismatch_table.upload(ex, t.data(), t.size());
		ex.sync();
	}
	read_filter_params p;
	p.stage_mask = (u32) params.filter_mask; p.stage_mask_hi = (u32) (params.filter_mask >> 32);
	p.homopolymer_length = params.homopolymer_length; p.min_read_through_distance = params.min_read_through_distance;
	p.max_overhang = 5; p.max_kmer_content = params.max_kmer_content; p.max_itd_length = params.max_itd_length;
	p.external_duplicate_marking = params.external_duplicate_marking;
	p.mismatch_table = mismatch_table.ptr(); p.table_n = table_n; p.table_k = table_k;
	return p;
}

// ------------------------------------------------------------------------------------------- duplicate marking
struct dup_key_fn { // pass 1: compact 12-byte keys, coalesced
	frag_view f; u64* k0; u32* k1; u8* participate;
	ARB_HD void operator()(u32 i) const {
		const dup_key k = duplicate_key(f, i);
		k0[i] = (u64) (u32) k.p1 | (u64) (u32) k.p2 << 32; k1[i] = (u32) k.c1 | (u32) k.c2 << 16;
		participate[i] = f.filter[i] == F_none;
	}
};
struct dup_key_ops {
	const u64* k0; const u32* k1;
	ARB_HD u64 hash(u32 i) const { dup_key k; k.p1 = (i32) (u32) k0[i]; k.p2 = (i32) (u32) (k0[i] >> 32); k.c1 = (u16) k1[i]; k.c2 = (u16) (k1[i] >> 16); return dup_hash(k); }
	ARB_HD bool equal(u32 a, u32 b) const { return k0[a] == k0[b] && k1[a] == k1[b]; }
};
struct dup_mark_fn {
	const u32* first; const u8* participate; u8* filter;
	ARB_HD void operator()(u32 i) const { if (participate[i] && first[i] != i) filter[i] = F_duplicates; }
};
struct dup_external_fn {
	frag_view f;
	ARB_HD void operator()(u32 i) const { if (f.filter[i] == F_none && (f.fflags[i] & FF_DUPLICATE)) f.filter[i] = F_duplicates; }
};

// ------------------------------------------------------------------------------------------- the cascade, two kernels
struct cascade_head_fn {
	read_filter_params p; frag_view f; annot_view an; u8* early; u32* queue; u32* n_queued;
	ARB_HD void operator()(u32 i) const {
		u8 e; bool more;
		const u8 label = classify_head(p, f, an, i, e, more);
		f.filter[i] = label; early[i] = e;
		if (more) queue[append_slot(n_queued)] = i;
	}
};
struct cascade_sequences_fn {
	read_filter_params p; frag_view f; annot_view an; const u32* queue;
	ARB_HD void operator()(u32 j, u32* scratch, u32 stride) const { const u32 i = queue[j]; f.filter[i] = classify_sequences(p, f, an, i, scratch, stride); }
};

// test hook: (mismatches, compared bases) of the two alignments the mismatch rule walks, on the packed reference and base by base
struct mismatch_probe_fn {
	frag_view f; annot_view packed, plain; u32* out; // 8 per fragment
	ARB_HD void operator()(u32 i) const {
		const u32 n = f.n_aln[i], a0 = f.idx(i, 0), a1 = f.idx(i, 1), a2 = f.idx(i, 2);
		const u32 y = n == 2 ? a1 : a2; const bool rc = n == 3 && f.fwd(a2) != f.fwd(a1);
		u32* o = out + (size_t) i * 8;
		count_mismatches(f, packed, a0, f.sq(a0), f.seq_len[a0], false, o[0], o[1]);
		count_mismatches(f, packed, y, f.sq(a1), f.seq_len[a1], rc, o[2], o[3]);
		count_mismatches(f, plain, a0, f.sq(a0), f.seq_len[a0], false, o[4], o[5]);
		count_mismatches(f, plain, y, f.sq(a1), f.seq_len[a1], rc, o[6], o[7]);
	}
};
void engine::probe_mismatch_counts(u32* out) {
	if (!annot.assembly4_ok) throw arb_error("the packed reference is disabled (characters outside the nt16 alphabet)");
	dbuf<u32> d((size_t) frags.n * 8);
	annot_view plain = annot.view(); plain.assembly4 = 0;
	mismatch_probe_fn fn = {frags.view(), annot.view(), plain, d.ptr()};
	for_each(ex, frags.n, fn);
	d.download(ex, out, (size_t) frags.n * 8);
}

struct count_labels_fn {
	const u8* filter; u32* counts;
	ARB_HD void operator()(u32 i) const { atomic_add_u32(&counts[filter[i] < F_COUNT ? filter[i] : 0], 1); }
};

void engine::run_read_filters() {
	if (!has_contigs || !has_annotation) throw arb_error("arb_run_read_filters: contigs and annotation must be set first");
	const u32 n = frags.n;
	const read_filter_params p = make_filter_params();
	frag_view f = frags.view();
	stage_timer t_all(ex);
	{
	stage_timer t_dup(ex);
	if (p.enabled(F_duplicates)) {
		if (p.external_duplicate_marking) {
			dup_external_fn fn = {f};
			for_each(ex, n, fn);
		} else {
			dbuf<u64> k0(n); dbuf<u32> k1(n), slot(n), first(n); dbuf<u8> part(n);
			dup_key_fn kf = {f, k0.ptr(), k1.ptr(), part.ptr()};
			for_each(ex, n, kf);
			dup_key_ops ops = {k0.ptr(), k1.ptr()};
			group_min_index(ex, table, n, ops, part.ptr(), slot.ptr(), first.ptr());
			dup_mark_fn mf = {first.ptr(), part.ptr(), f.filter};
			for_each(ex, n, mf);
			ex.sync();
		}
	}
	timings.duplicates_ms = t_dup.stop();
	}
	{
		dbuf<u32> queue(n), n_queued(1);
		n_queued.zero(ex, 1);
		stage_timer t_cls(ex);
		stage_timer t_head(ex);
		cascade_head_fn hf = {p, f, annot.view(), frags.early.ptr(), queue.ptr(), n_queued.ptr()};
		for_each(ex, n, hf);
		timings.cascade_head_ms = t_head.stop();
		u32 q = 0; n_queued.download(ex, &q, 1);
		stage_timer t_seq(ex);
		cascade_sequences_fn sf = {p, f, annot.view(), queue.ptr()};
		for_each_scratch<64>(ex, q, sf);
		timings.cascade_sequences_ms = t_seq.stop();
		timings.classify_ms = t_cls.stop();
		timings.cascade_queued = q;
		timings.cascade_algorithmic_bytes[0] = head_bytes;
		timings.cascade_algorithmic_bytes[1] = n ? (u64) ((double) sequence_bytes * q / n) : 0;
		timings.classify_algorithmic_bytes = timings.cascade_algorithmic_bytes[0] + timings.cascade_algorithmic_bytes[1];
	}
	timings.read_filters_ms = t_all.stop();
	ex.sync();
	filters_done = true;
}

void engine::get_fragment_filters(u8* filter_out, u8* early_out) {
	if (filter_out) frags.filter.download(ex, filter_out, frags.n);
	if (early_out) frags.early.download(ex, early_out, frags.n);
}

void engine::set_fragment_filters(const u8* filter) { frags.filter.upload(ex, filter, frags.n); ex.sync(); }

void engine::get_filter_counts(u32* counts) {
	label_counts.ensure(F_COUNT);
	label_counts.zero(ex, F_COUNT);
	count_labels_fn fn = {frags.filter.ptr(), label_counts.ptr()};
	for_each(ex, frags.n, fn);
	label_counts.download(ex, counts, F_COUNT);
}

} // namespace arb
