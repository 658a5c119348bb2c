// fusions.cu -- driver of candidate generation (see fusions_hd.h for the stage design and reference citations).
#include "engine.h"
#include "fusions_hd.h"

namespace arb {

static u32 bits_for(u32 n) { u32 b = 1; while (b < 32 && ((u64) 1 << b) < n) ++b; return b; }

struct fill_u32_fn { u32* p; u32 v; ARB_HD void operator()(u32 i) const { p[i] = v; } };

void engine::find_fusions(i32 max_mate_gap) {
	if (!has_annotation) throw arb_error("arb_find_fusions: annotation must be set first");
	const u32 n = frags.n;
	const u32 T = params.subsampling_threshold;
	frag_view f = frags.view();
	annot_view an = annot.view();
	stage_timer t_all(ex);

	// 1. records
	dbuf<u32> rec_off((size_t) n + 1);
	emit_count_fn ec = {f, rec_off.ptr()};
	for_each(ex, n, ec);
	exclusive_scan_u32(ex, rec_off.ptr(), rec_off.ptr(), n);
	u32 R = 0; rec_off.download(ex, &R, 1, n);
	dbuf<u32> r_gene1(R), r_gene2(R), r_contigs(R), r_meta(R), r_frag(R); dbuf<i32> r_bp1(R), r_bp2(R), r_an1(R), r_an2(R);
	record_view r = {r_gene1.ptr(), r_gene2.ptr(), r_contigs.ptr(), r_bp1.ptr(), r_bp2.ptr(), r_meta.ptr(), r_frag.ptr(), r_an1.ptr(), r_an2.ptr()};
	emit_fill_fn ef = {f, rec_off.ptr(), r};
	for_each(ex, n, ef);

	// 2. group records by candidate key; candidate id = rank of the group's first record
	dbuf<u32> first(R), slot(R), head((size_t) R + 1), cand_of(R), perm(R), tmp_k(R), tmp_v(R);
	record_key_ops kops = {r};
	group_min_index(ex, table, R, kops, (const u8*) NULL, slot.ptr(), first.ptr());
	head_flag_fn hf = {first.ptr(), head.ptr()};
	for_each(ex, R, hf);
	exclusive_scan_u32(ex, head.ptr(), head.ptr(), R);
	u32 C = 0; head.download(ex, &C, 1, R);
	group_id_fn gi = {first.ptr(), head.ptr(), cand_of.ptr(), perm.ptr()};
	for_each(ex, R, gi);

	// 3. contiguous, name-ordered segment per candidate
	radix_sort_pairs_u32(ex, cand_of.ptr(), perm.ptr(), tmp_k.ptr(), tmp_v.ptr(), R, bits_for(C));
	dbuf<u32> seg_off((size_t) C + 1);
	segment_offsets_fn so = {cand_of.ptr(), seg_off.ptr(), R, C};
	for_each(ex, R, so);

	cands.first_frag.ensure(C);
	first_fragment_fn ff = {r.frag, perm.ptr(), seg_off.ptr(), cands.first_frag.ptr()};
	for_each(ex, C, ff);

	// 4. pass A
	cands.n = C;
	cands.gene1.ensure(C); cands.gene2.ensure(C); cands.contig1.ensure(C); cands.contig2.ensure(C); cands.bp1.ensure(C); cands.bp2.ensure(C);
	cands.dir1.ensure(C); cands.dir2.ensure(C); cands.split_reads1.ensure(C); cands.split_reads2.ensure(C); cands.discordant_mates.ensure(C);
	cands.filter.ensure(C); cands.bits.ensure(C); cands.bits2.ensure(C); cands.anchor1.ensure(C); cands.anchor2.ensure(C); cands.evalue.ensure(C);
	cands.list1_off.ensure((size_t) C + 1); cands.list2_off.ensure((size_t) C + 1); cands.listd_off.ensure((size_t) C + 1);
	cand_out co = {cands.gene1.ptr(), cands.gene2.ptr(), cands.contig1.ptr(), cands.contig2.ptr(), cands.bp1.ptr(), cands.bp2.ptr(), cands.dir1.ptr(), cands.dir2.ptr(),
	               cands.split_reads1.ptr(), cands.split_reads2.ptr(), cands.discordant_mates.ptr(), cands.filter.ptr(), cands.bits.ptr(), cands.bits2.ptr(),
	               cands.anchor1.ptr(), cands.anchor2.ptr(), cands.evalue.ptr(), cands.list1_off.ptr(), cands.list2_off.ptr(), cands.listd_off.ptr()};
	dbuf<u8> kept(R);
	walk_a_fn wa = {r, perm.ptr(), seg_off.ptr(), co, kept.ptr(), T};
	for_each(ex, C, wa);
	// list sizes were written into the *_off arrays; keep copies of the counts, then scan in place
	dbuf<u32> n_list1((size_t) C + 1), n_list2((size_t) C + 1), n_listd((size_t) C + 1);
#ifdef ARB_DEVICE_BUILD
	ARB_CUDA_CHECK(cudaMemcpyAsync(n_list1.ptr(), cands.list1_off.ptr(), (size_t) C * 4, cudaMemcpyDeviceToDevice, ex.stream));
	ARB_CUDA_CHECK(cudaMemcpyAsync(n_list2.ptr(), cands.list2_off.ptr(), (size_t) C * 4, cudaMemcpyDeviceToDevice, ex.stream));
#else
	memcpy(n_list1.ptr(), cands.list1_off.ptr(), (size_t) C * 4); memcpy(n_list2.ptr(), cands.list2_off.ptr(), (size_t) C * 4);
#endif
	exclusive_scan_u32(ex, cands.list1_off.ptr(), cands.list1_off.ptr(), C);
	exclusive_scan_u32(ex, cands.list2_off.ptr(), cands.list2_off.ptr(), C);
	u32 L1 = 0, L2 = 0; cands.list1_off.download(ex, &L1, 1, C); cands.list2_off.download(ex, &L2, 1, C);
	cands.n_list1 = L1; cands.n_list2 = L2;
	cands.list1.ensure(L1); cands.list2.ensure(L2);
	fill_split_lists_fn fl = {r, perm.ptr(), seg_off.ptr(), kept.ptr(), cands.list1_off.ptr(), cands.list2_off.ptr(), cands.list1.ptr(), cands.list2.ptr()};
	for_each(ex, C, fl);

	// 5. discordant mates bucketed by (gene1, gene2, direction1, direction2), buckets in name order
	dbuf<u32> dflag((size_t) R + 1);
	flag_discordant_fn fd = {r, dflag.ptr()};
	for_each(ex, R, fd);
	exclusive_scan_u32(ex, dflag.ptr(), dflag.ptr(), R);
	u32 D = 0; dflag.download(ex, &D, 1, R);
	dbuf<u32> drec(D), bfirst(D), bslot(D), bhead((size_t) D + 1), bucket_of(D), bperm(D), btmp_k(D), btmp_v(D);
	compact_fn cf = {dflag.ptr(), NULL, drec.ptr(), R};
	for_each(ex, R, cf);
	hash_index bucket_table;
	bucket_key_ops bops = {r, drec.ptr()};
	group_min_index(ex, bucket_table, D, bops, (const u8*) NULL, bslot.ptr(), bfirst.ptr());
	head_flag_fn bhf = {bfirst.ptr(), bhead.ptr()};
	for_each(ex, D, bhf);
	exclusive_scan_u32(ex, bhead.ptr(), bhead.ptr(), D);
	u32 B = 0; bhead.download(ex, &B, 1, D);
	group_id_fn bgi = {bfirst.ptr(), bhead.ptr(), bucket_of.ptr(), bperm.ptr()};
	for_each(ex, D, bgi);
	radix_sort_pairs_u32(ex, bucket_of.ptr(), bperm.ptr(), btmp_k.ptr(), btmp_v.ptr(), D, bits_for(B));
	dbuf<u32> bseg_off((size_t) B + 1);
	segment_offsets_fn bso = {bucket_of.ptr(), bseg_off.ptr(), D, B};
	for_each(ex, D, bso);

	// 6. pass B (count, scan, fill) + canonical mate order
	cand_out cob = co; cob.n_list1 = n_list1.ptr(); cob.n_list2 = n_list2.ptr(); cob.n_listd = n_listd.ptr();
	dbuf<u32> need_swap(n); need_swap.zero(ex, n);
	walk_b_fn wb = {r, f, an, cob, bucket_table.view(), drec.ptr(), bhead.ptr(), bperm.ptr(), bseg_off.ptr(), NULL, NULL, need_swap.ptr(), max_mate_gap, T, 0};
	for_each(ex, C, wb);
	exclusive_scan_u32(ex, n_listd.ptr(), cands.listd_off.ptr(), C);
	u32 LD = 0; cands.listd_off.download(ex, &LD, 1, C);
	cands.n_listd = LD; cands.listd.ensure(LD);
	wb.listd_off = cands.listd_off.ptr(); wb.listd = cands.listd.ptr(); wb.fill = 1;
	for_each(ex, C, wb);
	swap_mates_fn sm = {f, need_swap.ptr(), frags.swapped.ptr()};
	for_each(ex, n, sm);

	// 7. pass C
	walk_c_fn wc = {f, an, cob, cands.list1_off.ptr(), cands.list2_off.ptr(), cands.listd_off.ptr(), cands.list1.ptr(), cands.list2.ptr(), cands.listd.ptr()};
	for_each(ex, C, wc);
	timings.find_fusions_ms = t_all.stop();
	ex.sync();
}

void engine::get_candidates(arb_candidates& o) {
	const u32 C = cands.n;
	cands.gene1.download(ex, o.gene1, C); cands.gene2.download(ex, o.gene2, C); cands.contig1.download(ex, o.contig1, C); cands.contig2.download(ex, o.contig2, C);
	cands.bp1.download(ex, o.breakpoint1, C); cands.bp2.download(ex, o.breakpoint2, C); cands.dir1.download(ex, o.direction1, C); cands.dir2.download(ex, o.direction2, C);
	cands.split_reads1.download(ex, o.split_reads1, C); cands.split_reads2.download(ex, o.split_reads2, C); cands.discordant_mates.download(ex, o.discordant_mates, C);
	cands.filter.download(ex, o.filter, C); cands.bits.download(ex, o.bits, C); cands.bits2.download(ex, o.bits2, C);
	cands.anchor1.download(ex, o.anchor_start1, C); cands.anchor2.download(ex, o.anchor_start2, C); cands.evalue.download(ex, o.evalue, C);
	cands.list1_off.download(ex, o.list1_off, (size_t) C + 1); cands.list2_off.download(ex, o.list2_off, (size_t) C + 1); cands.listd_off.download(ex, o.listd_off, (size_t) C + 1);
	cands.list1.download(ex, o.list1, cands.n_list1); cands.list2.download(ex, o.list2, cands.n_list2); cands.listd.download(ex, o.listd, cands.n_listd);
	o.n = C;
}

void engine::get_slot_swaps(u8* out) { frags.swapped.download(ex, out, frags.n); }
void engine::get_first_fragments(u32* out) { cands.first_frag.download(ex, out, cands.n); }

// installs a candidate table produced elsewhere (the merged shards of a multi-GPU run) in place of find_fusions' product
void engine::set_candidates(const arb_candidates& c) {
	const u32 C = c.n;
	cands.n = C;
	cands.gene1.upload(ex, c.gene1, C); cands.gene2.upload(ex, c.gene2, C); cands.contig1.upload(ex, c.contig1, C); cands.contig2.upload(ex, c.contig2, C);
	cands.bp1.upload(ex, c.breakpoint1, C); cands.bp2.upload(ex, c.breakpoint2, C); cands.dir1.upload(ex, c.direction1, C); cands.dir2.upload(ex, c.direction2, C);
	cands.split_reads1.upload(ex, c.split_reads1, C); cands.split_reads2.upload(ex, c.split_reads2, C); cands.discordant_mates.upload(ex, c.discordant_mates, C);
	cands.filter.upload(ex, c.filter, C); cands.bits.upload(ex, c.bits, C); cands.bits2.upload(ex, c.bits2, C);
	cands.anchor1.upload(ex, c.anchor_start1, C); cands.anchor2.upload(ex, c.anchor_start2, C); cands.evalue.upload(ex, c.evalue, C);
	cands.list1_off.upload(ex, c.list1_off, (size_t) C + 1); cands.list2_off.upload(ex, c.list2_off, (size_t) C + 1); cands.listd_off.upload(ex, c.listd_off, (size_t) C + 1);
	cands.n_list1 = c.list1_off[C]; cands.n_list2 = c.list2_off[C]; cands.n_listd = c.listd_off[C];
	cands.list1.upload(ex, c.list1, cands.n_list1); cands.list2.upload(ex, c.list2, cands.n_list2); cands.listd.upload(ex, c.listd, cands.n_listd);
	cands.first_frag.ensure(C); cands.first_frag.zero(ex, C);
	ex.sync();
}

// exchanges MATE1/MATE2 of the flagged fragments, as find_fusions does for listed discordant mates (fusions.cpp:414-421)
void engine::apply_slot_swaps(const u8* swapped) {
	const u32 n = frags.n;
	std::vector<u32> need(n);
	for (u32 i = 0; i < n; ++i) need[i] = swapped[i] ? 1u : 0u;
	dbuf<u32> need_swap; need_swap.upload(ex, need.data(), n);
	swap_mates_fn sm = {frags.view(), need_swap.ptr(), frags.swapped.ptr()};
	for_each(ex, n, sm);
	ex.sync();
}

} // namespace arb
