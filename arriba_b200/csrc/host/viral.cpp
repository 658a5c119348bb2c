// viral.cpp -- the two per-contig heuristics on viral contigs that sit between filter_viral_contigs and the fragment-length estimate
// (arriba.cpp:341-349): filter_top_expressed_viral_contigs (filter_top_expressed_viral_contigs.cpp:49-155) and
// filter_low_coverage_viral_contigs (filter_low_coverage_viral_contigs.cpp:12-53).
// Both decide PER CONTIG (from the mapped-read counts and the coverage windows collected at ingest, and from the gene sets of the fragments that join a
// viral and a host contig); the per-fragment part -- "any mate on a contig that was decided against" -- is two rules of the device cascade
// (read_filters.h, classify_head). The decisions travel as two more bits of the contig flags.
#include "pipeline.h"
#include <algorithm>
#include <set>

namespace arb { namespace host {

static u32 kmer12(const char* s) { // kmer_to_int (filter_mismappers.cpp:33-45): T=0 G=1 C=2, anything else 3
	u32 k = 0;
	for (int b = 0; b < 12; ++b) { const char c = s[b]; k = k << 2 | (c == 'T' ? 0u : c == 'G' ? 1u : c == 'C' ? 2u : 3u); }
	return k;
}

// filter_top_expressed_viral_contigs.cpp:22-47: at least a tenth of the distinct 12-mers of the smaller genome occur in the bigger one
static bool related_viral_strains(const char* a, size_t la, const char* b, size_t lb) {
	if (la > lb) { std::swap(a, b); std::swap(la, lb); }
	std::vector<u32> small;
	for (size_t i = 0; i + 12 <= la; ++i) small.push_back(kmer12(a + i));
	std::sort(small.begin(), small.end()); small.erase(std::unique(small.begin(), small.end()), small.end());
	std::vector<u8> seen(small.size(), 0);
	const unsigned min_shared = (unsigned) (small.size() / 10);
	unsigned shared = 0;
	for (size_t i = 0; i + 12 <= lb; ++i) {
		const u32 k = kmer12(b + i);
		std::vector<u32>::const_iterator it = std::lower_bound(small.begin(), small.end(), k);
		if (it != small.end() && *it == k && !seen[it - small.begin()]) { seen[it - small.begin()] = 1; if (++shared >= min_shared) return true; }
	}
	return false;
}

void viral_contig_decisions(pipeline& p) {
	refdata& ref = p.ref;
	const size_t nc = ref.contig_flags.size();
	bool any_viral = false;
	for (size_t c = 0; c < nc; ++c) { ref.contig_flags[c] &= (u8) ~(CF_VIRAL_LOW_EXPRESSION | CF_VIRAL_FOCAL_COVERAGE); if (ref.contig_flags[c] & CF_VIRAL) any_viral = true; }
	if (!any_viral) return;
	const u64 mask = p.opt.params.filter_mask;
	const frag_view f = p.frags.view();

	if (mask >> F_top_expressed_viral_contigs & 1) {
		const std::vector<u64>& reads = p.istats.mapped_viral_reads_by_contig;
		const size_t n = reads.size();
		std::vector<float> expression(n, 0.0f);
		for (size_t c = 0; c < n; ++c) if (ref.has_sequence((u32) c)) expression[c] = (float) (1.0 * reads[c] / ref.seq_len[c]);
		std::vector<u32> sorted(n);
		for (size_t c = 0; c < n; ++c) sorted[c] = (u32) c;
		std::sort(sorted.begin(), sorted.end(), [&](u32 x, u32 y) { return expression[x] != expression[y] ? expression[x] > expression[y] : x > y; });
		unsigned top_count = p.opt.top_viral_contigs, corrected = 0;
		for (size_t i = 1; i < n && expression[sorted[i]] > 0 && top_count > 0; ++i) { // related strains count as one
			++corrected;
			const u32 x = sorted[i], y = sorted[i - 1];
			if (!ref.has_sequence(x) || !ref.has_sequence(y) || !related_viral_strains(ref.sequence(x), ref.seq_len[x], ref.sequence(y), ref.seq_len[y])) --top_count;
		}
		if (corrected != 0) --corrected;
		const float min_expression = n ? expression[sorted[corrected]] : 0.0f;
		// viruses that integrate mostly between genes are kept unless they are among the least expressed ones (:97-130)
		const float min_intergenic_fraction = 0.33f;
		size_t spare = 50; if (spare > n) spare = n;
		const float min_expression_intergenic = n ? expression[sorted[n - spare]] : 0.0f;
		std::vector<std::set<u32> > sites(nc);
		for (u32 i = 0; i < f.n; ++i) {
			const u32 a = f.idx(i, MATE1), b = f.idx(i, f.n_aln[i] == 3 ? SUPPLEMENTARY : MATE2);
			int viral = -1, host = -1;
			if (ref.contig_flags[f.contig[a]] & CF_VIRAL) viral = (int) a; else if (ref.contig_flags[f.contig[a]] & CF_INTERESTING) host = (int) a;
			if (ref.contig_flags[f.contig[b]] & CF_VIRAL) viral = (int) b; else if (ref.contig_flags[f.contig[b]] & CF_INTERESTING) host = (int) b;
			if (viral >= 0 && host >= 0) sites[f.contig[viral]].insert(f.genes + f.genes_off[host], f.genes + f.genes_off[host] + f.genes_cnt[host]);
		}
		for (size_t c = 0; c < nc; ++c) {
			if (!(ref.contig_flags[c] & CF_VIRAL)) continue;
			unsigned intergenic = 0, genic = 0;
			for (std::set<u32>::const_iterator g = sites[c].begin(); g != sites[c].end(); ++g) if (ref.genes[*g].is_dummy) ++intergenic; else ++genic;
			const float fraction = intergenic > 0 ? (float) (1.0 * intergenic / (genic + intergenic)) : 0.0f;
			const float e = c < n ? expression[c] : 0.0f;
			if ((e == 0 || e < min_expression) && (fraction < min_intergenic_fraction || e == 0 || e < min_expression_intergenic)) ref.contig_flags[c] |= CF_VIRAL_LOW_EXPRESSION;
		}
	}

	if (mask >> F_low_coverage_viral_contigs & 1) {
		const float min_covered_fraction = p.opt.viral_contig_min_covered_fraction, min_covered_bases = 100;
		for (size_t c = 0; c < nc && c < p.coverage.coverage.size(); ++c) {
			if (!(ref.contig_flags[c] & CF_VIRAL)) continue;
			const std::vector<u16>& w = p.coverage.coverage[c];
			float average = 0;
			for (size_t k = 0; k < w.size(); ++k) average += w[k]; // float accumulation in window order, as the reference does
			average /= w.size();
			float sufficient = 0;
			for (size_t k = 0; k < w.size(); ++k) if (w[k] > 0.05 * average) sufficient++;
			if (sufficient / w.size() < min_covered_fraction || 20 /* COVERAGE_RESOLUTION */ * sufficient <= min_covered_bases) ref.contig_flags[c] |= CF_VIRAL_FOCAL_COVERAGE;
		}
	}
}

}} // namespace
