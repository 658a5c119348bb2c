// consensus.cu -- driver of the device-side pileups and consensus sequences of the surviving fusions (see consensus_hd.h).
#include <cstdlib>
#include "engine.h"
#include "consensus_hd.h"

namespace arb {

#ifdef ARB_DEVICE_BUILD
// one block per job; a block that runs out of jobs takes the next one of its stride (the grid is capped at what is resident)
template <u32 THREADS> __global__ void __launch_bounds__(THREADS) k_consensus(consensus_stage st, u32 n_jobs, u32 tiles) {
	extern __shared__ __align__(16) unsigned char consensus_shared[];
	pile_space s; s.carve(consensus_shared, tiles);
	const team_t t = {threadIdx.x, THREADS};
	for (u32 j = blockIdx.x; j < n_jobs; j += gridDim.x) { st.run(j, t, s); __syncthreads(); }
}
template <u32 THREADS> static void launch_consensus(const exec_ctx& ex, int device, const consensus_stage& st, u32 n_jobs, u32 tiles) {
	const size_t bytes = pile_space::bytes(tiles);
	ARB_CUDA_CHECK(cudaFuncSetAttribute(k_consensus<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes));
	int per_sm = 0, n_sm = 0;
	ARB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_consensus<THREADS>, (int) THREADS, bytes));
	ARB_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
	const u32 grid = std::min<u32>(n_jobs, (u32) std::max(1, per_sm) * (u32) n_sm * 4); // a few waves: jobs differ in cost by orders of magnitude
	k_consensus<THREADS><<<grid, THREADS, bytes, ex.stream>>>(st, n_jobs, tiles);
	ARB_CUDA_CHECK(cudaGetLastError());
	++stats().kernels;
}
#endif

static void run_jobs(const exec_ctx& ex, int device, const consensus_stage& st, u32 n_jobs, u32 tiles, bool wide) {
	if (n_jobs == 0) return;
#ifdef ARB_DEVICE_BUILD
	if (wide) launch_consensus<256>(ex, device, st, n_jobs, tiles); else launch_consensus<128>(ex, device, st, n_jobs, tiles);
#else
	(void) device; (void) wide;
	std::vector<unsigned long long> room(pile_space::bytes(tiles) / 8 + 1);
	pile_space s; s.carve(room.data(), tiles);
	const team_t t = {0, 1};
	for (u32 j = 0; j < n_jobs; ++j) st.run(j, t, s);
	++stats().kernels;
#endif
}

// rows: the candidates of the rows, in any order; afterwards arb_get_consensus hands out, per job 2 r + side, the consensus sequence, the positions of its
// characters and the clipped bases, and per row the number of non-template bases
void engine::build_consensus(const u32* rows, u32 R, arb_consensus_info& info) {
	memset(&info, 0, sizeof(info));
	cons_rows = R; cons_seq_bytes = cons_pos_count = cons_clip_bytes = 0;
	if (R == 0) return;
	if (!filters_done || cands.n == 0) throw arb_error("arb_build_consensus: no candidates are resident");
	if ((u64) R * 2 >= (1ull << 31)) throw arb_error("arb_build_consensus: too many rows in one call");
	stage_timer t_all(ex);
	const u32 J = 2 * R;
	u32 tiles_small = 32, tiles_wide = 256; // 1,024 / 8,192 positions per pileup: 21 KB / 162 KB of shared memory per block
	if (const char* e = getenv("ARB_CONSENSUS_TILES")) { tiles_small = (u32) std::max(1L, atol(e)); while (tiles_small & (tiles_small - 1)) ++tiles_small; tiles_wide = tiles_small * 4; } // test hook: force the retry and the host's share
	dbuf<u32> d_rows; d_rows.upload(ex, rows, R);
	dbuf<consensus_job_out> out(J);
	consensus_stage st;
	st.c = make_state_for_rows(); st.f = frags.view(); st.an = annot.view();
	st.l1o = cands.list1_off.ptr(); st.l1 = cands.list1.ptr(); st.l2o = cands.list2_off.ptr(); st.l2 = cands.list2.ptr(); st.ldo = cands.listd_off.ptr(); st.ld = cands.listd.ptr();
	st.rows = d_rows.ptr(); st.job_list = NULL; st.out = out.ptr(); st.launch = 0;
	// first launch: every job, small tables
	dbuf<u32> region((size_t) J + 1);
	{ consensus_capacity_fn cf = {st, tiles_small, region.ptr()}; for_each(ex, J, cf); }
	exclusive_scan_u32(ex, region.ptr(), region.ptr(), J);
	u32 total = 0; region.download(ex, &total, 1, J);
	if (total >= (1u << 30)) throw arb_error("arb_build_consensus: too many rows in one call");
	dbuf<char> chars((size_t) total * 2 + 16); dbuf<i32> pos((size_t) total + 4);
	st.region = region.ptr(); st.chars = chars.ptr(); st.pos = pos.ptr();
	run_jobs(ex, device, st, J, tiles_small, false);
	// second launch: what did not fit, four times the tiles
	dbuf<u32> flag((size_t) J + 1), retry(J), region2; dbuf<char> chars2; dbuf<i32> pos2;
	{ consensus_retry_fn rf = {out.ptr(), flag.ptr()}; for_each(ex, J, rf); }
	exclusive_scan_u32(ex, flag.ptr(), flag.ptr(), J);
	u32 n_retry = 0; flag.download(ex, &n_retry, 1, J);
	if (n_retry) {
		{ consensus_retry_gather_fn gf = {flag.ptr(), retry.ptr()}; for_each(ex, J, gf); }
		consensus_stage st2 = st; st2.job_list = retry.ptr(); st2.launch = 1;
		region2.alloc((size_t) n_retry + 1);
		{ consensus_capacity_fn cf = {st2, tiles_wide, region2.ptr()}; for_each(ex, n_retry, cf); }
		exclusive_scan_u32(ex, region2.ptr(), region2.ptr(), n_retry);
		u32 total2 = 0; region2.download(ex, &total2, 1, n_retry);
		if (total2 >= (1u << 30)) throw arb_error("arb_build_consensus: too many rows in one call");
		chars2.alloc((size_t) total2 * 2 + 16); pos2.alloc((size_t) total2 + 4);
		st2.region = region2.ptr(); st2.chars = chars2.ptr(); st2.pos = pos2.ptr();
		run_jobs(ex, device, st2, n_retry, tiles_wide, true);
	}
	// packed output
	cons_seq_off.ensure((size_t) J + 1); cons_pos_off.ensure((size_t) J + 1); cons_clip_off.ensure((size_t) J + 1); cons_verdict.ensure(J);
	{ consensus_lengths_fn lf = {out.ptr(), cons_seq_off.ptr(), cons_pos_off.ptr(), cons_clip_off.ptr(), cons_verdict.ptr()}; for_each(ex, J, lf); }
	exclusive_scan_u32(ex, cons_seq_off.ptr(), cons_seq_off.ptr(), J); exclusive_scan_u32(ex, cons_pos_off.ptr(), cons_pos_off.ptr(), J); exclusive_scan_u32(ex, cons_clip_off.ptr(), cons_clip_off.ptr(), J);
	u32 n_seq = 0, n_pos = 0, n_clip = 0;
	cons_seq_off.download(ex, &n_seq, 1, J); cons_pos_off.download(ex, &n_pos, 1, J); cons_clip_off.download(ex, &n_clip, 1, J);
	cons_seq.ensure((size_t) n_seq + 1); cons_pos.ensure((size_t) n_pos + 1); cons_clip.ensure((size_t) n_clip + 1);
	{
		consensus_pack_fn pf; pf.out = out.ptr(); pf.chars[0] = chars.ptr(); pf.chars[1] = chars2.ptr(); pf.pos[0] = pos.ptr(); pf.pos[1] = pos2.ptr();
		pf.seq_off = cons_seq_off.ptr(); pf.pos_off = cons_pos_off.ptr(); pf.clip_off = cons_clip_off.ptr(); pf.seq_out = cons_seq.ptr(); pf.pos_out = cons_pos.ptr(); pf.clip_out = cons_clip.ptr();
		for_each(ex, J, pf);
	}
	cons_non_template.ensure(R);
	{ non_template_fn nf = {frags.view(), d_rows.ptr(), st.l1o, st.l1, st.l2o, st.l2, cons_non_template.ptr()}; for_each(ex, R, nf); }
	cons_seq_bytes = n_seq; cons_pos_count = n_pos; cons_clip_bytes = n_clip;
	timings.consensus_ms += t_all.stop();
	ex.sync();
	info.n_rows = R; info.seq_bytes = n_seq; info.pos_count = n_pos; info.clip_bytes = n_clip; info.retried_jobs = n_retry;
}

void engine::get_consensus(u32* seq_off, u32* pos_off, u32* clip_off, u8* verdict, u32* non_template, char* seq, i32* pos, char* clip) {
	const u32 J = 2 * cons_rows;
	if (J == 0) return;
	cons_seq_off.download(ex, seq_off, (size_t) J + 1); cons_pos_off.download(ex, pos_off, (size_t) J + 1); cons_clip_off.download(ex, clip_off, (size_t) J + 1);
	cons_verdict.download(ex, verdict, J); cons_non_template.download(ex, non_template, cons_rows);
	cons_seq.download(ex, seq, cons_seq_bytes); cons_pos.download(ex, pos, cons_pos_count); cons_clip.download(ex, clip, cons_clip_bytes);
}

} // namespace arb
