// annotate.cu -- driver of the device-side annotation of the resident fragments (see annotate_hd.h for the rules and the reference citations).
//
//   arb_annotate_pass1   : after arb_push_chunk_begin (columns of ingest resident; the alignment flags travel with this call) and arb_set_annotation (genes of
//                          the GTF): strands from the library type, gene sets / exonic flags / strands per alignment, then the clusters of breakpoints
//                          that got no gene (sort by (contig, position), neighbour test, scan) = the dummy genes the caller appends to its gene table
//   arb_annotate_pass2   : after arb_set_annotation with the dummy genes appended: remaining breakpoints get their dummy gene, the gene sets are laid out
//                          as the CSR columns of the resident table, which is complete afterwards (no arb_push_chunk_end)
//   arb_get_annotation_columns : the columns the host-side event logic reads (alignment flags, gene sets)
#include "engine.h"
#include "annotate_hd.h"

namespace arb {

enum { ANNOT_CAP_PASS1 = 32, ANNOT_CAP_PASS2 = 65535 }; // genes under one alignment pass 1 can hold / the 16-bit gene count column can express (loud error beyond)

static gene_sets_view sets_view(engine& e) {
	gene_sets_view v = {e.annot_rows.ptr(), e.annot_cnt.ptr(), e.annot_pool.ptr(), e.annot_ctl.ptr(), e.annot_pool_cap, e.annot_ctl.ptr() + 1};
	return v;
}

static void check_annotation_errors(engine& e, const char* pass, int cap) {
	u32 ctl[2] = {0, 0};
	e.annot_ctl.download(e.ex, ctl, 2);
	if (ctl[1] & 1u) throw arb_error(std::string("annotation ") + pass + ": more than " + std::to_string(cap) + " overlapping genes under one alignment are not supported");
	if (ctl[1] & 2u) throw arb_error(std::string("annotation ") + pass + ": gene set pool exhausted");
}

u32 engine::annotate_pass1(const u8* aflags, int strandedness) {
	if (!push_open) throw arb_error("arb_annotate_pass1 without arb_push_chunk_begin");
	if (!has_annotation) throw arb_error("arb_annotate_pass1: annotation must be set first");
	const u32 n = frags.n; const size_t A = 3 * (size_t) n;
	{ // the alignment flags as ingest left them, on the copy stream behind the other columns; then the compute stream may read the table
		stage_timer t_h2d(copy_ex);
		frags.aflags.upload(copy_ex, aflags, A);
		timings.h2d_ms += t_h2d.stop(); timings.h2d_bytes += A;
	}
	stage_timer t_all(ex);
	annot_rows.ensure(A * gene_sets_view::ROW); annot_cnt.ensure(A);
	annot_pool_cap = (u32) std::min<u64>(0xFFFFFFF0ull, (u64) A / 2 + (1u << 20)); // sets of more than ROW genes are rare
	annot_pool.ensure(annot_pool_cap);
	annot_ctl.ensure(2); annot_ctl.zero(ex, 2);
	frags.genes_off.ensure(A); frags.genes_cnt.ensure(A);
	annotate_pass1_fn<ANNOT_CAP_PASS1> p1 = {annot.view(), frags.view(), sets_view(*this), strandedness};
	for_each(ex, n, p1);
	check_annotation_errors(*this, "pass 1", ANNOT_CAP_PASS1);
	// breakpoints without a gene, sorted by (contig, position)
	dbuf<u32> uoff((size_t) n + 1);
	unmapped_count_fn uc = {frags.view(), annot_cnt.ptr(), uoff.ptr()};
	for_each(ex, n, uc);
	exclusive_scan_u32(ex, uoff.ptr(), uoff.ptr(), n);
	u32 U = 0; uoff.download(ex, &U, 1, n);
	n_dummy = 0;
	if (U) {
		dbuf<u32> ucontig(U), upos(U), tk(U), tv(U);
		unmapped_fill_fn uf = {frags.view(), annot_cnt.ptr(), uoff.ptr(), ucontig.ptr(), upos.ptr()};
		for_each(ex, n, uf);
		u32 cbits = 1; while (cbits < 16 && (1u << cbits) < annot.n_contigs) ++cbits;
		radix_sort_pairs_u32(ex, upos.ptr(), ucontig.ptr(), tk.ptr(), tv.ptr(), U, 32);
		radix_sort_pairs_u32(ex, ucontig.ptr(), upos.ptr(), tk.ptr(), tv.ptr(), U, cbits);
		dbuf<u32> brk((size_t) U + 1), brk_scan((size_t) U + 1);
		dummy_break_fn db = {annot.view(), ucontig.ptr(), upos.ptr(), brk.ptr()};
		for_each(ex, U, db);
		exclusive_scan_u32(ex, brk.ptr(), brk_scan.ptr(), U);
		brk_scan.download(ex, &n_dummy, 1, U);
		dummy_contig.ensure(n_dummy); dummy_start.ensure(n_dummy); dummy_end.ensure(n_dummy);
		dummy_emit_fn de = {ucontig.ptr(), upos.ptr(), brk.ptr(), brk_scan.ptr(), U, dummy_contig.ptr(), dummy_start.ptr(), dummy_end.ptr()};
		for_each(ex, U, de);
	}
	timings.annotate_ms = t_all.stop();
	return n_dummy;
}

void engine::get_dummy_genes(u16* contig, i32* start, i32* end) {
	dummy_contig.download(ex, contig, n_dummy); dummy_start.download(ex, start, n_dummy); dummy_end.download(ex, end, n_dummy);
}

u64 engine::annotate_pass2() {
	if (!push_open) throw arb_error("arb_annotate_pass2 without arb_annotate_pass1");
	const u32 n = frags.n; const size_t A = 3 * (size_t) n;
	if (A > 0xFFFFFFF0ull) throw arb_error("chunk too large");
	stage_timer t_all(ex);
	annotate_pass2_fn p2 = {annot.view(), frags.view(), sets_view(*this)};
	for_each(ex, n, p2);
	check_annotation_errors(*this, "pass 2", ANNOT_CAP_PASS2);
	dbuf<u32> off(A + 1);
	gene_count_fn gc = {annot_cnt.ptr(), off.ptr()};
	for_each(ex, (u32) A, gc);
	exclusive_scan_u32(ex, off.ptr(), off.ptr(), (u32) A);
	u32 G = 0; off.download(ex, &G, 1, A);
	frags.genes.ensure((size_t) G + 1);
	gene_fill_fn gf = {sets_view(*this), off.ptr(), frags.view()};
	for_each(ex, (u32) A, gf);
	n_gene_entries = G;
	timings.annotate_ms += t_all.stop();
	ex.sync();
	annot_rows.release(); annot_pool.release();
	finish_push(G);
	return G;
}

void engine::get_annotation_columns(u8* aflags, u32* genes_off, u16* genes_cnt, u32* genes) {
	const size_t A = 3 * (size_t) frags.n;
	frags.aflags.download(ex, aflags, A); frags.genes_off.download(ex, genes_off, A); frags.genes_cnt.download(ex, genes_cnt, A); frags.genes.download(ex, genes, n_gene_entries);
}

} // namespace arb
