// shard.cpp -- one sample on several GPUs.
//
// Fragments are partitioned by the unordered pair of contigs their two ends lie on. Everything the device does up to and including candidate generation
// only relates fragments of the same pair: the duplicate key holds both contigs (filter_duplicates.cpp:25-42), a candidate's key holds the contigs of both
// breakpoints (fusions.cpp:253-300) and discordant mates are attached through gene pairs, which imply the contigs. Two exchange steps remain:
//   labels      after the read-level cascade, because estimate_fragment_length (read_stats.cpp:17-107) looks at the first fragments in NAME order;
//   candidates  after find_fusions, because everything from merge_adjacent_fusions on relates candidates across contig pairs.
// The transport is the caller's (torch.distributed / NCCL all-gather in bench.py and the tests): a rank exports one blob per exchange and imports the
// blobs of all ranks. After the second exchange every rank holds the complete fragment table and the merged candidate table, numbered as a single
// device would have numbered them (by first insertion), so the rest of the run and the output are byte-identical to the single-GPU run.
#include <algorithm>
#include <cstring>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include "pipeline.h"

namespace arb { namespace host {

static void check(arb_ctx* ctx, int rc, const char* what) { if (rc != 0) throw std::runtime_error(std::string(what) + ": " + arb_last_error(ctx)); }

static inline u32 pair_key(u32 a, u32 b) { return a < b ? a << 16 | b : b << 16 | a; }

// ------------------------------------------------------------------------------------------- partition
void pipeline::set_shard(int rank, int world) {
	if (world < 1 || rank < 0 || rank >= world) throw std::runtime_error("invalid shard rank / world size");
	if (ctx && frags_on_device) throw std::runtime_error("the shard must be chosen before the upload step");
	shard_rank = rank; shard_world = world;
	shard_members.assign((size_t) world, std::vector<u32>());
	if (world == 1) return;
	const size_t N = frags.n;
	// contig pair the candidates of a fragment live in, and the pair its duplicate key lives in; both must end up on the same rank
	std::vector<u32> key(N);
	std::vector<u32> keys; // distinct keys, later sorted
	std::vector<std::pair<u32, u32> > links;
	for (size_t i = 0; i < N; ++i) {
		const u32 c0 = frags.contig[i], c1 = frags.contig[N + i], c2 = frags.contig[2 * N + i];
		if (frags.n_aln[i] == 3) { key[i] = pair_key(c1, c2); const u32 dup = pair_key(c0, c2); if (dup != key[i]) links.push_back(std::make_pair(key[i], dup)); }
		else key[i] = pair_key(c0, c1);
	}
	keys = key; for (size_t k = 0; k < links.size(); ++k) keys.push_back(links[k].second);
	std::sort(keys.begin(), keys.end()); keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
	auto id_of = [&](u32 k) { return (u32) (std::lower_bound(keys.begin(), keys.end(), k) - keys.begin()); };
	std::vector<u32> parent(keys.size()); std::iota(parent.begin(), parent.end(), 0u);
	auto find = [&](u32 x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
	for (size_t k = 0; k < links.size(); ++k) { const u32 a = find(id_of(links[k].first)), b = find(id_of(links[k].second)); if (a != b) parent[std::max(a, b)] = std::min(a, b); }
	std::vector<u64> weight(keys.size(), 0);
	std::vector<u32> comp(N);
	for (size_t i = 0; i < N; ++i) { comp[i] = find(id_of(key[i])); ++weight[comp[i]]; }
	// longest-processing-time assignment of the components (a plain hash would skew: intra-chromosomal pairs dominate)
	std::vector<u32> by_weight;
	for (u32 c = 0; c < keys.size(); ++c) if (weight[c] > 0) by_weight.push_back(c);
	std::sort(by_weight.begin(), by_weight.end(), [&](u32 a, u32 b) { return weight[a] != weight[b] ? weight[a] > weight[b] : a < b; });
	std::vector<u64> load((size_t) world, 0); std::vector<int> owner(keys.size(), 0);
	for (size_t k = 0; k < by_weight.size(); ++k) {
		int best = 0; for (int r = 1; r < world; ++r) if (load[r] < load[best]) best = r;
		owner[by_weight[k]] = best; load[best] += weight[by_weight[k]];
	}
	for (size_t i = 0; i < N; ++i) shard_members[owner[comp[i]]].push_back((u32) i);
	// the shard's own fragment table: same columns, pools re-packed
	const std::vector<u32>& mine = shard_members[rank];
	const size_t n = mine.size();
	fragment_table& l = local;
	l.n = (u32) n;
	l.n_aln.resize(n); l.fflags.resize(n); l.filter.resize(n);
	l.contig.resize(3 * n); l.start.resize(3 * n); l.end.resize(3 * n); l.aflags.resize(3 * n); l.cigar_off.resize(3 * n); l.cigar_cnt.resize(3 * n);
	l.seq_off.resize(2 * n); l.seq_len.resize(2 * n); l.genes_off.resize(3 * n); l.genes_cnt.resize(3 * n);
	u64 n_cigar = 0, n_seq = 0, n_genes = 0;
	for (size_t j = 0; j < n; ++j) {
		const size_t i = mine[j];
		for (u32 s = 0; s < 3; ++s) { n_cigar += frags.cigar_cnt[s * N + i]; n_genes += frags.genes_cnt[s * N + i]; if (s < 2) n_seq += ((frags.seq_len[s * N + i] + 1) / 2 + 15) / 16; }
	}
	l.cigar.resize(n_cigar + 1); l.genes.resize(n_genes + 1); l.seq.resize(n_seq * 16 + 16);
	u64 c_at = 0, s_at = 0, g_at = 0;
	for (size_t j = 0; j < n; ++j) {
		const size_t i = mine[j];
		l.n_aln[j] = frags.n_aln[i]; l.fflags[j] = frags.fflags[i]; l.filter[j] = frags.filter[i];
		for (u32 s = 0; s < 3; ++s) {
			const size_t x = s * N + i, y = s * n + j;
			l.contig[y] = frags.contig[x]; l.start[y] = frags.start[x]; l.end[y] = frags.end[x]; l.aflags[y] = frags.aflags[x];
			l.cigar_off[y] = (u32) c_at; l.cigar_cnt[y] = frags.cigar_cnt[x];
			memcpy(&l.cigar[c_at], &frags.cigar[frags.cigar_off[x]], 4ull * frags.cigar_cnt[x]); c_at += frags.cigar_cnt[x];
			l.genes_off[y] = (u32) g_at; l.genes_cnt[y] = frags.genes_cnt[x];
			memcpy(&l.genes[g_at], &frags.genes[frags.genes_off[x]], 4ull * frags.genes_cnt[x]); g_at += frags.genes_cnt[x];
			if (s < 2) {
				const u64 units = ((frags.seq_len[x] + 1) / 2 + 15) / 16;
				l.seq_off[y] = (u32) s_at; l.seq_len[y] = frags.seq_len[x];
				memcpy(&l.seq[s_at * 16], &frags.seq[(size_t) frags.seq_off[x] * 16], units * 16); s_at += units;
			}
		}
	}
	l.cigar[c_at] = 0; l.genes[g_at] = 0; memset(&l.seq[s_at * 16], 0, 16);
}

// ------------------------------------------------------------------------------------------- exchange
namespace {
struct writer { std::vector<u8>& b; template <class T> void put(const T* p, size_t n) { const size_t at = b.size(); b.resize(at + n * sizeof(T)); if (n) memcpy(&b[at], p, n * sizeof(T)); } template <class T> void one(T v) { put(&v, 1); } };
struct reader { const u8* p; const u8* end; template <class T> void get(T* out, size_t n) { if ((size_t) (end - p) < n * sizeof(T)) throw std::runtime_error("truncated shard blob"); if (n) memcpy(out, p, n * sizeof(T)); p += n * sizeof(T); } template <class T> T one() { T v; get(&v, 1); return v; } };
}

void pipeline::export_shard(int what, const void** blob, u64* bytes) {
	if (shard_world <= 1) throw std::runtime_error("the run is not sharded");
	export_blob.clear();
	writer w = {export_blob};
	const u32 n = local.n;
	w.one<u32>((u32) what); w.one<u32>((u32) shard_rank); w.one<u32>(n);
	if (what == ARB_EXCHANGE_LABELS) {
		w.put(local_labels.data(), n); w.put(local_early.data(), n);
	} else if (what == ARB_EXCHANGE_CANDIDATES) {
		uint32_t C; uint64_t n1, n2, nd;
		check(ctx, arb_candidates_size(ctx, &C, &n1, &n2, &nd), "arb_candidates_size");
		event_table e; // scratch image of the shard's table
		e.gene1.resize(C); e.gene2.resize(C); e.contig1.resize(C); e.contig2.resize(C); e.bp1.resize(C); e.bp2.resize(C); e.dir1.resize(C); e.dir2.resize(C);
		e.split_reads1.resize(C); e.split_reads2.resize(C); e.discordant_mates.resize(C); e.filter.resize(C); e.bits.resize(C); e.bits2.resize(C);
		e.anchor1.resize(C); e.anchor2.resize(C); e.evalue.resize(C);
		e.list1_off.resize((size_t) C + 1); e.list2_off.resize((size_t) C + 1); e.listd_off.resize((size_t) C + 1); e.list1.resize(n1 + 1); e.list2.resize(n2 + 1); e.listd.resize(nd + 1);
		arb_candidates c;
		c.n = C; c.gene1 = e.gene1.data(); c.gene2 = e.gene2.data(); c.contig1 = e.contig1.data(); c.contig2 = e.contig2.data(); c.breakpoint1 = e.bp1.data(); c.breakpoint2 = e.bp2.data();
		c.direction1 = e.dir1.data(); c.direction2 = e.dir2.data(); c.split_reads1 = e.split_reads1.data(); c.split_reads2 = e.split_reads2.data(); c.discordant_mates = e.discordant_mates.data();
		c.filter = e.filter.data(); c.bits = e.bits.data(); c.bits2 = e.bits2.data(); c.anchor_start1 = e.anchor1.data(); c.anchor_start2 = e.anchor2.data(); c.evalue = e.evalue.data();
		c.list1_off = e.list1_off.data(); c.list2_off = e.list2_off.data(); c.listd_off = e.listd_off.data(); c.list1 = e.list1.data(); c.list2 = e.list2.data(); c.listd = e.listd.data();
		check(ctx, arb_get_candidates(ctx, &c), "arb_get_candidates");
		std::vector<u32> first(C);
		check(ctx, arb_get_candidate_first_fragments(ctx, first.data()), "arb_get_candidate_first_fragments");
		std::vector<u8> swapped(n), lab(n);
		check(ctx, arb_get_slot_swaps(ctx, swapped.data()), "arb_get_slot_swaps");
		check(ctx, arb_get_fragment_filters(ctx, lab.data(), NULL), "arb_get_fragment_filters");
		// fragment indices leave the rank as GLOBAL name ranks
		const std::vector<u32>& mine = shard_members[shard_rank];
		for (u64 k = 0; k < n1; ++k) e.list1[k] = mine[e.list1[k]];
		for (u64 k = 0; k < n2; ++k) e.list2[k] = mine[e.list2[k]];
		for (u64 k = 0; k < nd; ++k) e.listd[k] = mine[e.listd[k]];
		for (u32 k = 0; k < C; ++k) first[k] = mine[first[k]];
		w.one<u32>(C); w.one<u64>(n1); w.one<u64>(n2); w.one<u64>(nd);
		w.put(e.gene1.data(), C); w.put(e.gene2.data(), C); w.put(e.contig1.data(), C); w.put(e.contig2.data(), C); w.put(e.bp1.data(), C); w.put(e.bp2.data(), C);
		w.put(e.dir1.data(), C); w.put(e.dir2.data(), C); w.put(e.split_reads1.data(), C); w.put(e.split_reads2.data(), C); w.put(e.discordant_mates.data(), C);
		w.put(e.filter.data(), C); w.put(e.bits.data(), C); w.put(e.bits2.data(), C); w.put(e.anchor1.data(), C); w.put(e.anchor2.data(), C); w.put(e.evalue.data(), C);
		w.put(e.list1_off.data(), (size_t) C + 1); w.put(e.list2_off.data(), (size_t) C + 1); w.put(e.listd_off.data(), (size_t) C + 1);
		w.put(e.list1.data(), n1); w.put(e.list2.data(), n2); w.put(e.listd.data(), nd);
		w.put(first.data(), C); w.put(swapped.data(), n); w.put(lab.data(), n);
	} else throw std::runtime_error("unknown exchange");
	*blob = export_blob.data(); *bytes = export_blob.size();
}

void pipeline::import_shards(int what, const void* const* blobs, const u64* bytes, u32 n_blobs) {
	if (shard_world <= 1) throw std::runtime_error("the run is not sharded");
	if ((int) n_blobs != shard_world) throw std::runtime_error("expected one blob per rank");
	const size_t N = frags.n;
	if (what == ARB_EXCHANGE_LABELS) {
		labels.assign(N, 0); early.assign(N, 0);
		for (u32 r = 0; r < n_blobs; ++r) {
			reader in = {(const u8*) blobs[r], (const u8*) blobs[r] + bytes[r]};
			if (in.one<u32>() != (u32) what || in.one<u32>() != r) throw std::runtime_error("shard blobs out of order");
			const std::vector<u32>& members = shard_members[r];
			const u32 n = in.one<u32>();
			if (n != members.size()) throw std::runtime_error("shard sizes disagree between ranks");
			std::vector<u8> l(n), e(n); in.get(l.data(), n); in.get(e.data(), n);
			for (u32 j = 0; j < n; ++j) { labels[members[j]] = l[j]; early[members[j]] = e[j]; }
		}
		say_read_filter_counts();
		return;
	}
	if (what != ARB_EXCHANGE_CANDIDATES) throw std::runtime_error("unknown exchange");
	// ---- gather the tables of all ranks
	struct part { event_table e; std::vector<u32> first; };
	std::vector<part> parts(n_blobs);
	std::vector<u8> swapped(N, 0);
	u64 C_total = 0, n1_total = 0, n2_total = 0, nd_total = 0;
	for (u32 r = 0; r < n_blobs; ++r) {
		reader in = {(const u8*) blobs[r], (const u8*) blobs[r] + bytes[r]};
		if (in.one<u32>() != (u32) what || in.one<u32>() != r) throw std::runtime_error("shard blobs out of order");
		const std::vector<u32>& members = shard_members[r];
		const u32 n = in.one<u32>();
		if (n != members.size()) throw std::runtime_error("shard sizes disagree between ranks");
		event_table& e = parts[r].e;
		const u32 C = in.one<u32>(); const u64 n1 = in.one<u64>(), n2 = in.one<u64>(), nd = in.one<u64>();
		e.n = C;
		e.gene1.resize(C); e.gene2.resize(C); e.contig1.resize(C); e.contig2.resize(C); e.bp1.resize(C); e.bp2.resize(C); e.dir1.resize(C); e.dir2.resize(C);
		e.split_reads1.resize(C); e.split_reads2.resize(C); e.discordant_mates.resize(C); e.filter.resize(C); e.bits.resize(C); e.bits2.resize(C);
		e.anchor1.resize(C); e.anchor2.resize(C); e.evalue.resize(C);
		e.list1_off.resize((size_t) C + 1); e.list2_off.resize((size_t) C + 1); e.listd_off.resize((size_t) C + 1); e.list1.resize(n1); e.list2.resize(n2); e.listd.resize(nd);
		in.get(e.gene1.data(), C); in.get(e.gene2.data(), C); in.get(e.contig1.data(), C); in.get(e.contig2.data(), C); in.get(e.bp1.data(), C); in.get(e.bp2.data(), C);
		in.get(e.dir1.data(), C); in.get(e.dir2.data(), C); in.get(e.split_reads1.data(), C); in.get(e.split_reads2.data(), C); in.get(e.discordant_mates.data(), C);
		in.get(e.filter.data(), C); in.get(e.bits.data(), C); in.get(e.bits2.data(), C); in.get(e.anchor1.data(), C); in.get(e.anchor2.data(), C); in.get(e.evalue.data(), C);
		in.get(e.list1_off.data(), (size_t) C + 1); in.get(e.list2_off.data(), (size_t) C + 1); in.get(e.listd_off.data(), (size_t) C + 1);
		in.get(e.list1.data(), n1); in.get(e.list2.data(), n2); in.get(e.listd.data(), nd);
		parts[r].first.resize(C); in.get(parts[r].first.data(), C);
		std::vector<u8> sw(n), lab(n); in.get(sw.data(), n); in.get(lab.data(), n);
		for (u32 j = 0; j < n; ++j) { swapped[members[j]] = sw[j]; labels[members[j]] = lab[j]; }
		C_total += C; n1_total += n1; n2_total += n2; nd_total += nd;
	}
	if (C_total > 0xFFFFFFFFull) throw std::runtime_error("more than 2^32 candidates");
	// ---- candidate ids of a single device: first insertion = (first fragment in name order, order within that fragment == local id)
	std::vector<std::pair<u32, u32> > src; src.reserve(C_total); // (rank, local id)
	for (u32 r = 0; r < n_blobs; ++r) for (u32 k = 0; k < parts[r].e.n; ++k) src.push_back(std::make_pair(r, k));
	std::stable_sort(src.begin(), src.end(), [&](const std::pair<u32, u32>& a, const std::pair<u32, u32>& b) {
		const u32 fa = parts[a.first].first[a.second], fb = parts[b.first].first[b.second];
		if (fa != fb) return fa < fb;
		return a.second < b.second; // same fragment => same rank
	});
	event_table m;
	const u32 C = (u32) C_total;
	m.n = C;
	m.gene1.resize(C); m.gene2.resize(C); m.contig1.resize(C); m.contig2.resize(C); m.bp1.resize(C); m.bp2.resize(C); m.dir1.resize(C); m.dir2.resize(C);
	m.split_reads1.resize(C); m.split_reads2.resize(C); m.discordant_mates.resize(C); m.filter.resize(C); m.bits.resize(C); m.bits2.resize(C);
	m.anchor1.resize(C); m.anchor2.resize(C); m.evalue.resize(C);
	m.list1_off.assign((size_t) C + 1, 0); m.list2_off.assign((size_t) C + 1, 0); m.listd_off.assign((size_t) C + 1, 0);
	m.list1.resize(n1_total + 1); m.list2.resize(n2_total + 1); m.listd.resize(nd_total + 1);
	for (u32 k = 0; k < C; ++k) {
		const event_table& e = parts[src[k].first].e; const u32 j = src[k].second;
		m.gene1[k] = e.gene1[j]; m.gene2[k] = e.gene2[j]; m.contig1[k] = e.contig1[j]; m.contig2[k] = e.contig2[j]; m.bp1[k] = e.bp1[j]; m.bp2[k] = e.bp2[j]; m.dir1[k] = e.dir1[j]; m.dir2[k] = e.dir2[j];
		m.split_reads1[k] = e.split_reads1[j]; m.split_reads2[k] = e.split_reads2[j]; m.discordant_mates[k] = e.discordant_mates[j]; m.filter[k] = e.filter[j]; m.bits[k] = e.bits[j]; m.bits2[k] = e.bits2[j];
		m.anchor1[k] = e.anchor1[j]; m.anchor2[k] = e.anchor2[j]; m.evalue[k] = e.evalue[j];
		m.list1_off[k + 1] = m.list1_off[k] + (e.list1_off[j + 1] - e.list1_off[j]); m.list2_off[k + 1] = m.list2_off[k] + (e.list2_off[j + 1] - e.list2_off[j]); m.listd_off[k + 1] = m.listd_off[k] + (e.listd_off[j + 1] - e.listd_off[j]);
		if (e.list1_off[j + 1] > e.list1_off[j]) memcpy(&m.list1[m.list1_off[k]], &e.list1[e.list1_off[j]], 4ull * (e.list1_off[j + 1] - e.list1_off[j]));
		if (e.list2_off[j + 1] > e.list2_off[j]) memcpy(&m.list2[m.list2_off[k]], &e.list2[e.list2_off[j]], 4ull * (e.list2_off[j + 1] - e.list2_off[j]));
		if (e.listd_off[j + 1] > e.listd_off[j]) memcpy(&m.listd[m.listd_off[k]], &e.listd[e.listd_off[j]], 4ull * (e.listd_off[j + 1] - e.listd_off[j]));
	}
	// ---- from here on this rank continues like a single device: whole fragment table, labels, mate order, merged candidates
	shard_world = 1; // upload() and the event stages address the complete table again
	{ fragment_table empty; std::swap(local, empty); }
	frags_on_device = false;
	upload();
	check(ctx, arb_set_fragment_filters(ctx, labels.data()), "arb_set_fragment_filters");
	check(ctx, arb_apply_slot_swaps(ctx, swapped.data()), "arb_apply_slot_swaps");
	arb_candidates c;
	c.n = C; c.gene1 = m.gene1.data(); c.gene2 = m.gene2.data(); c.contig1 = m.contig1.data(); c.contig2 = m.contig2.data(); c.breakpoint1 = m.bp1.data(); c.breakpoint2 = m.bp2.data();
	c.direction1 = m.dir1.data(); c.direction2 = m.dir2.data(); c.split_reads1 = m.split_reads1.data(); c.split_reads2 = m.split_reads2.data(); c.discordant_mates = m.discordant_mates.data();
	c.filter = m.filter.data(); c.bits = m.bits.data(); c.bits2 = m.bits2.data(); c.anchor_start1 = m.anchor1.data(); c.anchor_start2 = m.anchor2.data(); c.evalue = m.evalue.data();
	c.list1_off = m.list1_off.data(); c.list2_off = m.list2_off.data(); c.listd_off = m.listd_off.data(); c.list1 = m.list1.data(); c.list2 = m.list2.data(); c.listd = m.listd.data();
	check(ctx, arb_set_candidates(ctx, &c), "arb_set_candidates");
}

}} // namespace
