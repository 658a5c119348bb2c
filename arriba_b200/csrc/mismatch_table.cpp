// mismatch_table.cpp -- host-side decision table for the `mismatches` filter.
//
// filter_mismatches.cpp:55-99 turns (mismatches k, compared bases n) into a keep/discard decision through a binomial
// p-value (double arithmetic, result rounded to float) and a random-hit model that uses x87 long double. Both depend on
// small integers only, so the decision is tabulated once on the host -- with the same C++ types, operation order and
// libm as the reference build -- and the kernel (read_filters.h: too_many_mismatches) only counts (n, k).
// All counters are 32-bit unsigned on purpose: k > n wraps exactly like the reference's `alignment_length - mismatches`.
#include "mismatch_table.h"
#include <cmath>

namespace arb {

static double binomial_coefficient(const unsigned int k, const unsigned int n) {
	double c = 1;
	for (unsigned int i = n - k + 1; i <= n; ++i) c *= i;
	for (unsigned int i = 1; i <= k; ++i) c /= i;
	return c;
}

bool mismatch_decision(const unsigned int k, const unsigned int n, const float p, const unsigned long genome_size, const float cutoff) {
	const float binomial_p = binomial_coefficient(k, n) * std::pow(p, k) * std::pow(1 - p, n - k); // double product, rounded to float
	if (binomial_p < cutoff) return true;
	if (k == 0) return false;
	const long double permutations = std::pow(4, n - k);
	if (genome_size >= permutations) return true;
	return (1 - std::pow(1 - genome_size / permutations, binomial_coefficient(k, n))) > 0.01;
}

std::vector<uint8_t> build_mismatch_table(unsigned int table_n, unsigned int table_k, float p, unsigned long genome_size, float cutoff) {
	std::vector<uint8_t> t((size_t) table_n * table_k);
	for (unsigned int n = 0; n < table_n; ++n)
		for (unsigned int k = 0; k < table_k; ++k)
			t[(size_t) n * table_k + k] = mismatch_decision(k, n, p, genome_size, cutoff);
	return t;
}

} // namespace arb
