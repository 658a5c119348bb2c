// shard.cpp -- one sample on several GPUs, host side: the balanced assignment of contig pairs to the parts, and the split
// re-alignment stage. The device side and the design are in csrc/exchange.cu; the transport belongs to the launcher (arriba_b200/sharded.py).
//
// Everything find_fusions does relates fragments of the same unordered contig pair only: a candidate's key holds the contigs of both breakpoints
// (fusions.cpp:253-300) and discordant mates are attached through gene pairs, which imply the contigs. A part that emits the breakpoints of a set of contig
// pairs therefore finds exactly the candidates of those pairs. Pairs that share duplicates (filter_duplicates.cpp:25-42 keys a split read by the contigs of
// MATE1 and the supplementary) are kept together as well, so that the same assignment also closes the read-level cascade.
#include <algorithm>
#include <cstring>
#include <map>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <thread>
#include "pipeline.h"

namespace arb { namespace host {

static void check(arb_ctx* ctx, int rc, const char* what) { if (rc != 0) throw std::runtime_error(std::string(what) + ": " + arb_last_error(ctx)); }
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static inline u32 pair_key(u32 a, u32 b) { return a < b ? a << 16 | b : b << 16 | a; }

// contig pairs (ascending keys) and the part that owns each: connected components of pairs linked by duplicates, assigned heaviest first to the lightest part
void pipeline::work_partition(int parts) {
	if (parts < 1 || parts > 255) throw std::runtime_error("invalid number of parts");
	const size_t N = frags.n;
	const int T = std::max(1, threads);
	std::vector<std::map<u32, u64> > weight_of(T); std::vector<std::map<u64, bool> > links_of(T); // a sample has a few thousand distinct contig pairs at most
	std::vector<std::thread> pool;
	for (int t = 0; t < T; ++t) pool.emplace_back([&, t]() {
		std::map<u32, u64>& w = weight_of[t]; std::map<u64, bool>& l = links_of[t];
		u32 last_key = 0xFFFFFFFFu; u64* last_count = NULL; u64 last_link = ~(u64) 0;
		for (size_t i = N * t / T; i < N * (t + 1) / T; ++i) {
			const u32 c0 = frags.contig[i], c1 = frags.contig[N + i], c2 = frags.contig[2 * N + i];
			u32 key;
			if (frags.n_aln[i] == 3) { key = pair_key(c1, c2); const u32 dup = pair_key(c0, c2); if (dup != key) { const u64 link = (u64) key << 32 | dup; if (link != last_link) { l[link] = true; last_link = link; } } }
			else key = pair_key(c0, c1);
			if (key != last_key) { last_count = &w[key]; last_key = key; }
			++*last_count;
		}
	});
	for (size_t t = 0; t < pool.size(); ++t) pool[t].join();
	std::map<u32, u64> weight; std::map<u64, bool> links;
	for (int t = 0; t < T; ++t) { for (std::map<u32, u64>::iterator k = weight_of[t].begin(); k != weight_of[t].end(); ++k) weight[k->first] += k->second; links.insert(links_of[t].begin(), links_of[t].end()); }
	for (std::map<u64, bool>::iterator k = links.begin(); k != links.end(); ++k) weight[(u32) k->first]; // a duplicate key's pair may hold no candidates itself
	std::vector<u32>& keys = partition_keys; std::vector<u8>& owner = partition_owner;
	keys.clear(); std::vector<u64> w;
	for (std::map<u32, u64>::iterator k = weight.begin(); k != weight.end(); ++k) { keys.push_back(k->first); w.push_back(k->second); }
	auto id_of = [&](u32 k) { return (u32) (std::lower_bound(keys.begin(), keys.end(), k) - keys.begin()); };
	std::vector<u32> parent(keys.size()); std::iota(parent.begin(), parent.end(), 0u);
	auto find = [&](u32 x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
	for (std::map<u64, bool>::iterator k = links.begin(); k != links.end(); ++k) { const u32 a = find(id_of((u32) (k->first >> 32))), b = find(id_of((u32) k->first)); if (a != b) parent[std::max(a, b)] = std::min(a, b); }
	std::vector<u64> comp_weight(keys.size(), 0);
	for (u32 k = 0; k < keys.size(); ++k) comp_weight[find(k)] += w[k];
	std::vector<u32> by_weight;
	for (u32 c = 0; c < keys.size(); ++c) if (find(c) == c) by_weight.push_back(c);
	std::sort(by_weight.begin(), by_weight.end(), [&](u32 a, u32 b) { return comp_weight[a] != comp_weight[b] ? comp_weight[a] > comp_weight[b] : a < b; });
	std::vector<u64> load((size_t) parts, 0); std::vector<int> comp_owner(keys.size(), 0);
	for (size_t k = 0; k < by_weight.size(); ++k) { // a plain hash would skew: intra-chromosomal pairs dominate
		int best = 0; for (int r = 1; r < parts; ++r) if (load[r] < load[best]) best = r;
		comp_owner[by_weight[k]] = best; load[best] += comp_weight[by_weight[k]];
	}
	owner.resize(keys.size());
	for (u32 k = 0; k < keys.size(); ++k) owner[k] = (u8) comp_owner[find(k)];
}

// ------------------------------------------------------------------------------------------- the re-alignment stage in two halves
// A multi-GPU launcher runs the event chain up to filter_mismappers, lets every part re-align its share of the work items (arb_filter_mismappers_part on
// each part's context, verdicts combined by an all-reduce) and resumes the chain. mismappers_begin returns false when the stage is switched off.
bool pipeline::mismappers_begin() {
	events_until(EV_MISMAPPERS - 1);
	if (!((opt.params.filter_mask >> F_mismappers) & 1)) return false;
	t_mismappers_begin = now_s();
	check(ctx, arb_set_fragment_filters(ctx, labels.data()), "arb_set_fragment_filters");
	push_candidate_state();
	return true;
}

void pipeline::mismappers_end() {
	if (events_done != EV_MISMAPPERS - 1) throw std::runtime_error("mismappers_end without mismappers_begin");
	if ((opt.params.filter_mask >> F_mismappers) & 1) {
		pull_candidate_state();
		check(ctx, arb_get_fragment_filters(ctx, labels.data(), NULL), "arb_get_fragment_filters");
		log_remaining("Re-aligning chimeric reads to filter fusions with >=80% mis-mappers");
		t_events[EV_MISMAPPERS] = now_s() - t_mismappers_begin;
	}
	events_done = EV_MISMAPPERS;
}

}} // namespace
