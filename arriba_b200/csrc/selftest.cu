// selftest.cu -- C entry points that run the device primitives on caller data (used by tests/test_prims.py to check the
// hand-written scan / radix sort / hash group-by kernels against numpy). Not part of the reference-facing interface.
#include "prims.h"

using namespace arb;

struct u32_key_ops { const u32* k; ARB_HD u64 hash(u32 i) const { u64 h = k[i] * 0x9E3779B97F4A7C15ULL; return h ^ (h >> 29); } ARB_HD bool equal(u32 a, u32 b) const { return k[a] == k[b]; } };

extern "C" {

uint64_t arb_selftest_pool_size_class(uint64_t bytes) { return pool_size_class((size_t) bytes); }


int arb_selftest_scan(const uint32_t* in, uint32_t* out /* n+1 */, uint32_t n) {
	try {
		scratch_set scratch; exec_ctx ex; ex.scratch = &scratch;
		dbuf<u32> d((size_t) n + 1);
		d.upload(ex, in, n);
		exclusive_scan_u32(ex, d.ptr(), d.ptr(), n);
		d.download(ex, out, (size_t) n + 1);
	} catch (const std::exception&) { return 1; }
	return 0;
}

int arb_selftest_sort(uint32_t* keys, uint32_t* vals, uint32_t n, uint32_t bits) {
	try {
		scratch_set scratch; exec_ctx ex; ex.scratch = &scratch;
		dbuf<u32> k(n), v(n), kt(n), vt(n);
		k.upload(ex, keys, n); v.upload(ex, vals, n);
		radix_sort_pairs_u32(ex, k.ptr(), v.ptr(), kt.ptr(), vt.ptr(), n, bits);
		k.download(ex, keys, n); v.download(ex, vals, n);
	} catch (const std::exception&) { return 1; }
	return 0;
}

int arb_selftest_group(const uint32_t* keys, uint32_t* first, uint32_t n) {
	try {
		scratch_set scratch; exec_ctx ex; ex.scratch = &scratch;
		dbuf<u32> k(n), slot(n), f(n);
		k.upload(ex, keys, n);
		hash_index t;
		u32_key_ops ops = {k.ptr()};
		group_min_index(ex, t, n, ops, (const u8*) NULL, slot.ptr(), f.ptr());
		f.download(ex, first, n);
	} catch (const std::exception&) { return 1; }
	return 0;
}

}
