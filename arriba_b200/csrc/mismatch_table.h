// mismatch_table.h -- host-side decision table for the `mismatches` filter.
//
// filter_mismatches.cpp:55-99 turns (mismatches k, compared bases n) into a keep/discard decision through a binomial
// p-value (double arithmetic, result rounded to float) and a random-hit model that uses x87 long double. Both depend on
// small integers only, so the decision is tabulated once on the host -- with the same C++ types, operation order and
// libm as the reference build -- and the kernel (read_filters.h: too_many_mismatches) only counts (n, k).
// All counters are 32-bit unsigned on purpose: k > n wraps exactly like the reference's `alignment_length - mismatches`.
#pragma once
#include <vector>
#include <stdint.h>
#include <stddef.h>

namespace arb {
bool mismatch_decision(unsigned int k, unsigned int n, float p, unsigned long genome_size, float cutoff);
// table[n * table_k + k]; compiled by the host C++ compiler only (mismatch_table.cpp), never by the device compiler
std::vector<uint8_t> build_mismatch_table(unsigned int table_n, unsigned int table_k, float p, unsigned long genome_size, float cutoff);
}
