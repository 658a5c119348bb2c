// ingest.h -- chimeric BAM ingest: BGZF/BAM decoding on host threads into structure-of-arrays fragment columns
// (arb_soa_chunk layout, fragments in name order), plus the by-products the later stages need (mapped read count,
// coverage windows). Behavioural contract: read_chimeric_alignments (read_chimeric_alignments.cpp:560-773) and helpers
// (:19-558), mark_multimappers (:792-802), coverage_t (read_stats.cpp:148-306). Only STAR "WithinBAM" input (-x) is
// handled; Chimeric.out.sam (-c), SAM text and CRAM are not.
#pragma once
#include <string>
#include <vector>
#include <cstdlib>
#include <memory>
#include <utility>
#include <new>
#include "refdata.h"
struct arb_ctx;

namespace arb { namespace host {

struct coverage_windows { // 20 bp windows (read_stats.hpp:14)
	std::vector<std::vector<u16> > coverage;
	std::vector<std::vector<u8> > starts, ends;
	void resize(const refdata& ref);
	bool fragment_starts_here(u32 contig, i32 start, i32 end) const;
	bool fragment_ends_here(u32 contig, i32 start, i32 end) const;
	int get_coverage(u32 contig, i32 position, u32 direction) const;
};

// Large host blocks are recycled inside the process: the first touch of fresh pages costs ~2 us per 4 KiB page on the measured hosts (more than filling
// them), and a process that runs sample after sample would pay it for gigabytes of columns every time. Blocks of 1 MiB and more go back to a free list per
// size class (steps of 25 %) instead of the C library; host_block_trim() (arb_release_host_memory) hands them to the system.
void* host_block_get(size_t bytes, size_t& granted);
void release_worker_cache(); // the ingest keeps the workers' pools of a finished sample for the next one (ingest.cpp)
void host_block_put(void* p, size_t granted);
void host_block_trim();
// where fresh blocks come from (default: malloc). The CUDA library installs cudaHostAlloc / cudaFreeHost (capi.cu); a NULL result falls back to malloc.
void set_host_block_backend(void* (*alloc)(size_t), void (*release)(void*));

// vector whose resize() leaves new elements uninitialised: the large columns are first touched (and zeroed where needed) by the threads that fill them
template <class T> struct default_init_allocator {
	typedef T value_type;
	template <class U> struct rebind { typedef default_init_allocator<U> other; };
	default_init_allocator() {}
	template <class U> default_init_allocator(const default_init_allocator<U>&) {}
	static size_t class_of(size_t bytes) { size_t c = (size_t) 1 << 20; while (c < bytes) c += c / 4; return c; }
	T* allocate(size_t n) {
		const size_t bytes = n * sizeof(T);
		void* p;
		if (bytes >= ((size_t) 1 << 20)) { size_t granted; p = host_block_get(class_of(bytes), granted); }
		else p = malloc(bytes ? bytes : 1);
		if (!p) throw std::bad_alloc();
		return (T*) p;
	}
	void deallocate(T* p, size_t n) { const size_t bytes = n * sizeof(T); if (bytes >= ((size_t) 1 << 20)) host_block_put(p, class_of(bytes)); else free(p); }
	template <class U> void construct(U* p) { ::new ((void*) p) U; }
	template <class U, class A0, class... A> void construct(U* p, A0&& a0, A&&... a) { ::new ((void*) p) U(std::forward<A0>(a0), std::forward<A>(a)...); }
	template <class U> bool operator==(const default_init_allocator<U>&) const { return true; }
	template <class U> bool operator!=(const default_init_allocator<U>&) const { return false; }
};
template <class T> using column = std::vector<T, default_init_allocator<T> >;

struct fragment_table { // host image of arb_soa_chunk + names
	u32 n;
	column<u8> n_aln, fflags, filter, aflags;
	column<u16> contig, cigar_cnt, seq_len, genes_cnt;
	column<i32> start, end;
	column<u32> cigar_off, seq_off, genes_off, cigar, genes;
	column<u8> seq;
	column<char> names; column<u64> name_off; // "<qname>,<HI>[ITD]" per fragment, name order
	fragment_table(): n(0) {}
	frag_view view();
	std::string name(u32 i) const { return std::string(names.data() + name_off[i], names.data() + name_off[i + 1]); }
};

struct ingest_stats {
	u64 mapped_reads; std::vector<u64> mapped_viral_reads_by_contig; u64 malformed, missing_hi_tag, records; bool no_chimeric_reads;
	double t_inflate, t_parse, t_finalize;
};

struct ingest_options { bool external_duplicate_marking; u32 max_itd_length; std::string interesting_contigs, viral_contigs; int threads; struct ::arb_ctx* scan_ctx /* device context that finds the records of every chunk (arb_bam_scan) */; };

// reads the BAM, fills `out` (name order, slots normalised, multimappers marked); throws std::runtime_error on fatal input errors
void read_chimeric_alignments(const std::string& bam_path, refdata& ref, const ingest_options& opt, fragment_table& out, coverage_windows& coverage, ingest_stats& stats);

}} // namespace
