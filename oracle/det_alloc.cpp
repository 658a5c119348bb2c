// oracle/det_alloc.cpp -- TEST INFRASTRUCTURE (linked only into oracle/_ref/arriba).
//
// The reference orders gene sets and exon sets by POINTER VALUE (annotation_set_t over gene_t / exon_t,
// common.hpp:128-143,153,177) and the pointers are std::list nodes (common.hpp:144), so the order of a
// multi-gene set -- and through it the insertion order into fusions_t (fusions.cpp:248-252) and the gene
// picked by estimate_fragment_length (read_stats.cpp:33) -- depends on what glibc malloc happened to
// return. With ARB_DET_ALLOC=1 this replacement of the global operator new serves allocations of exactly
// the two list-node sizes from a never-reusing bump arena, which makes pointer order == creation order
// (GTF order for annotated genes/exons, then dummy genes in creation order). The product defines the
// same order. Without the variable, everything goes to malloc and the binary is the plain reference.
#include <cstdlib>
#include <cstdint>
#include <list>
#include <new>
#include <sys/mman.h>
#include "common.hpp"

namespace {
const size_t GENE_NODE = sizeof(std::_List_node<gene_annotation_record_t>);
const size_t EXON_NODE = sizeof(std::_List_node<exon_annotation_record_t>);
const size_t ARENA_BYTES = (size_t) 1 << 36; // virtual reservation only
char* arena = NULL; size_t used = 0; int enabled = -1;
inline bool active() {
	if (enabled < 0) {
		const char* e = getenv("ARB_DET_ALLOC");
		enabled = e && e[0] == '1';
		if (enabled) {
			arena = (char*) mmap(NULL, ARENA_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
			if (arena == MAP_FAILED) { arena = NULL; enabled = 0; }
		}
	}
	return enabled;
}
}

void* operator new(size_t n) {
	if ((n == GENE_NODE || n == EXON_NODE) && active() && used + n + 16 < ARENA_BYTES) {
		void* p = arena + used;
		used += (n + 15) & ~(size_t) 15;
		return p;
	}
	void* p = malloc(n ? n : 1);
	if (!p) throw std::bad_alloc();
	return p;
}
void operator delete(void* p) noexcept {
	if (arena && (char*) p >= arena && (char*) p < arena + ARENA_BYTES) return;
	free(p);
}
void operator delete(void* p, size_t) noexcept { operator delete(p); }
void* operator new[](size_t n) { return operator new(n); }
void operator delete[](void* p) noexcept { operator delete(p); }
void operator delete[](void* p, size_t) noexcept { operator delete(p); }
