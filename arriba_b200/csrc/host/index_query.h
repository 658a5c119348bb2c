// index_query.h -- host-side front end of query_index (annot_hd.h) for loops that run once per fragment or per output row.
// The result sets must be able to hold pathological loci (thousands of overlapping records), but a set of that capacity is tens of kilobytes of stack per call,
// three of them per range query; touched (and probed) on every call, they cost more than the query. Nearly every locus holds a handful of records: the query is
// made with a small set first and repeated with the large one only where the small one overflowed. Same code path, same result.
#pragma once
#include "../annot_hd.h"
#include <stdexcept>
#include <string>

namespace arb { namespace host {

template <int BIG, class F> __attribute__((noinline)) void index_query_big(const region_index_view& ix, u32 contig, i32 start, i32 end, F& use, const char* overflow_message) {
	idset<BIG> big;
	query_index(ix, contig, start, end, big);
	if (big.overflow) throw std::runtime_error(overflow_message);
	use((const u32*) big.v, big.n);
}

// use(ids, n): ascending item ids of the regions that the point (start == end) or range query returns (annotation.t.hpp:55-100)
template <int BIG, class F> inline void index_query(const region_index_view& ix, u32 contig, i32 start, i32 end, F use, const char* overflow_message) {
	idset<48> small;
	query_index(ix, contig, start, end, small);
	if (!small.overflow) { use((const u32*) small.v, small.n); return; }
	index_query_big<BIG>(ix, contig, start, end, use, overflow_message);
}

}} // namespace
