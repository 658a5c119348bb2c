// oracle/shim/shim.cpp -- TEST INFRASTRUCTURE (oracle build only; never linked into the product).
//
// Implementation of the htslib subset declared in sam.h / bgzf.h / cram.h over system zlib.
// Written from the SAM/BAM specification (SAMv1 section 4.2: BAM record layout; section 4.1:
// BGZF = concatenated gzip members) -- no htslib source was available or consulted.
// parity unpinned at the htslib boundary: the reference ships no test vectors for BAM decoding;
// tests/test_ingest.py and tests/test_z_bam_corpus.py pin this reader against BAMs from two independent spec-level writers (tools/synth.cpp, tests/bamtools.py).
#include <zlib.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include "sam.h"
#include "bgzf.h"
#include "cram.h"

extern "C" {

const char seq_nt16_str[] = "=ACMGRSVTWYHKDBN";

// ---------------------------------------------------------------- gzip / BGZF stream
struct BGZF {
	FILE* f;
	z_stream zs;
	bool compressed;
	bool zs_active;
	bool eof;
	unsigned char* in;
	size_t in_cap;
};

BGZF* bgzf_open(const char* path, const char* mode) {
	(void) mode;
	FILE* f = fopen(path, "rb");
	if (!f) return NULL;
	BGZF* fp = (BGZF*) calloc(1, sizeof(BGZF));
	fp->f = f;
	fp->in_cap = 1 << 20;
	fp->in = (unsigned char*) malloc(fp->in_cap);
	setvbuf(f, NULL, _IONBF, 0);
	size_t n = fread(fp->in, 1, fp->in_cap, f);
	fp->compressed = n >= 2 && fp->in[0] == 0x1f && fp->in[1] == 0x8b;
	memset(&fp->zs, 0, sizeof(fp->zs));
	fp->zs.next_in = fp->in;
	fp->zs.avail_in = n;
	if (fp->compressed) {
		if (inflateInit2(&fp->zs, 15 + 32) != Z_OK) { fclose(f); free(fp->in); free(fp); return NULL; }
		fp->zs_active = true;
	}
	return fp;
}

static bool bgzf_refill(BGZF* fp) {
	if (fp->zs.avail_in > 0) return true;
	size_t n = fread(fp->in, 1, fp->in_cap, fp->f);
	fp->zs.next_in = fp->in;
	fp->zs.avail_in = n;
	return n > 0;
}

ssize_t bgzf_read(BGZF* fp, void* data, size_t length) {
	unsigned char* out = (unsigned char*) data;
	size_t done = 0;
	if (!fp->compressed) {
		while (done < length) {
			if (!bgzf_refill(fp)) break;
			size_t k = length - done < fp->zs.avail_in ? length - done : fp->zs.avail_in;
			memcpy(out + done, fp->zs.next_in, k);
			fp->zs.next_in += k; fp->zs.avail_in -= k; done += k;
		}
		return done;
	}
	while (done < length && !fp->eof) {
		if (!bgzf_refill(fp)) { fp->eof = true; break; }
		fp->zs.next_out = out + done;
		fp->zs.avail_out = length - done;
		int r = inflate(&fp->zs, Z_NO_FLUSH);
		done = length - fp->zs.avail_out;
		if (r == Z_STREAM_END) {
			// next gzip member (BGZF block), if any
			if (!bgzf_refill(fp)) { fp->eof = true; break; }
			if (inflateReset(&fp->zs) != Z_OK) return -1;
		} else if (r != Z_OK && r != Z_BUF_ERROR) {
			return -1;
		}
	}
	return done;
}

int bgzf_close(BGZF* fp) {
	if (!fp) return 0;
	if (fp->zs_active) inflateEnd(&fp->zs);
	fclose(fp->f);
	free(fp->in);
	free(fp);
	return 0;
}

// ---------------------------------------------------------------- BAM
static inline uint32_t le32(const unsigned char* p) { return p[0] | p[1] << 8 | p[2] << 16 | (uint32_t) p[3] << 24; }
static inline uint16_t le16(const unsigned char* p) { return p[0] | p[1] << 8; }

samFile* sam_open(const char* path, const char* mode) {
	(void) mode;
	BGZF* b = bgzf_open(path, "rb");
	if (!b) return NULL;
	samFile* fp = (samFile*) calloc(1, sizeof(samFile));
	fp->is_bin = 1;
	fp->is_bgzf = 1;
	fp->is_cram = 0;
	fp->fp.bgzf = b;
	return fp;
}

int sam_close(samFile* fp) {
	if (!fp) return 0;
	bgzf_close(fp->fp.bgzf);
	free(fp);
	return 0;
}

int hts_set_threads(samFile* fp, int n) { (void) fp; (void) n; return 0; } // single-threaded decode

sam_hdr_t* sam_hdr_read(samFile* fp) {
	BGZF* b = fp->fp.bgzf;
	unsigned char buf[8];
	if (bgzf_read(b, buf, 4) != 4 || memcmp(buf, "BAM\1", 4) != 0) return NULL;
	if (bgzf_read(b, buf, 4) != 4) return NULL;
	uint32_t l_text = le32(buf);
	char* text = (char*) malloc(l_text + 1);
	if (bgzf_read(b, text, l_text) != (ssize_t) l_text) { free(text); return NULL; }
	free(text);
	if (bgzf_read(b, buf, 4) != 4) return NULL;
	sam_hdr_t* h = (sam_hdr_t*) calloc(1, sizeof(sam_hdr_t));
	h->n_targets = le32(buf);
	h->target_len = (uint32_t*) calloc(h->n_targets ? h->n_targets : 1, sizeof(uint32_t));
	h->target_name = (char**) calloc(h->n_targets ? h->n_targets : 1, sizeof(char*));
	for (int i = 0; i < h->n_targets; ++i) {
		if (bgzf_read(b, buf, 4) != 4) return NULL;
		uint32_t l_name = le32(buf);
		h->target_name[i] = (char*) malloc(l_name + 1);
		if (bgzf_read(b, h->target_name[i], l_name) != (ssize_t) l_name) return NULL;
		h->target_name[i][l_name] = '\0';
		if (bgzf_read(b, buf, 4) != 4) return NULL;
		h->target_len[i] = le32(buf);
	}
	return h;
}

void sam_hdr_destroy(sam_hdr_t* h) {
	if (!h) return;
	for (int i = 0; i < h->n_targets; ++i) free(h->target_name[i]);
	free(h->target_name);
	free(h->target_len);
	free(h);
}

bam1_t* bam_init1(void) { return (bam1_t*) calloc(1, sizeof(bam1_t)); }

void bam_destroy1(bam1_t* b) {
	if (!b) return;
	free(b->data);
	free(b);
}

int sam_read1(samFile* fp, sam_hdr_t* h, bam1_t* b) {
	(void) h;
	BGZF* z = fp->fp.bgzf;
	unsigned char x[36];
	ssize_t r = bgzf_read(z, x, 4);
	if (r == 0) return -1; // EOF
	if (r != 4) return -2;
	uint32_t block_size = le32(x);
	if (block_size < 32) return -2;
	if (bgzf_read(z, x, 32) != 32) return -2;
	bam1_core_t* c = &b->core;
	c->tid = (int32_t) le32(x);
	c->pos = (int32_t) le32(x + 4);
	uint32_t l_read_name = x[8];
	c->qual = x[9];
	c->bin = le16(x + 10);
	c->n_cigar = le16(x + 12);
	c->flag = le16(x + 14);
	c->l_qseq = (int32_t) le32(x + 16);
	c->mtid = (int32_t) le32(x + 20);
	c->mpos = (int32_t) le32(x + 24);
	c->isize = (int32_t) le32(x + 28);
	// pad the read name with NULs so that the CIGAR array is 4-byte aligned
	uint32_t extranul = (4 - (l_read_name & 3)) & 3;
	c->l_extranul = extranul;
	c->l_qname = l_read_name + extranul;
	uint32_t rest = block_size - 32;
	uint32_t need = rest + extranul;
	if (b->m_data < need) {
		b->m_data = need + 32;
		b->m_data += b->m_data >> 1;
		b->data = (uint8_t*) realloc(b->data, b->m_data);
		if (!b->data) return -3;
	}
	if (bgzf_read(z, b->data, l_read_name) != (ssize_t) l_read_name) return -2;
	for (uint32_t i = 0; i < extranul; ++i) b->data[l_read_name + i] = '\0';
	if (bgzf_read(z, b->data + c->l_qname, rest - l_read_name) != (ssize_t) (rest - l_read_name)) return -2;
	b->l_data = need;
	return 0;
}

static inline int aux_type_size(uint8_t t) {
	switch (t) {
		case 'A': case 'c': case 'C': return 1;
		case 's': case 'S': return 2;
		case 'i': case 'I': case 'f': return 4;
		case 'd': return 8;
		default: return 0;
	}
}

uint8_t* bam_aux_get(const bam1_t* b, const char tag[2]) {
	const uint8_t* s = bam_get_aux(b);
	const uint8_t* end = b->data + b->l_data;
	while (s + 3 <= end) {
		bool hit = s[0] == (uint8_t) tag[0] && s[1] == (uint8_t) tag[1];
		const uint8_t* val = s + 2; // points at the type byte
		uint8_t t = val[0];
		const uint8_t* next;
		if (t == 'Z' || t == 'H') {
			next = val + 1;
			while (next < end && *next) ++next;
			++next;
		} else if (t == 'B') {
			if (val + 6 > end) return NULL;
			int sz = aux_type_size(val[1]);
			uint32_t n = le32(val + 2);
			next = val + 6 + (size_t) sz * n;
		} else {
			int sz = aux_type_size(t);
			if (sz == 0) return NULL;
			next = val + 1 + sz;
		}
		if (hit) return (uint8_t*) val;
		s = next;
	}
	return NULL;
}

int64_t bam_aux2i(const uint8_t* s) {
	switch (s[0]) {
		case 'c': return (int8_t) s[1];
		case 'C': return s[1];
		case 's': return (int16_t) le16(s + 1);
		case 'S': return le16(s + 1);
		case 'i': return (int32_t) le32(s + 1);
		case 'I': return le32(s + 1);
		default: return 0;
	}
}

int64_t bam_cigar2qlen(int n_cigar, const uint32_t* cigar) {
	int64_t l = 0;
	for (int k = 0; k < n_cigar; ++k)
		if (bam_cigar_type(bam_cigar_op(cigar[k])) & 1)
			l += bam_cigar_oplen(cigar[k]);
	return l;
}

hts_pos_t bam_cigar2rlen(int n_cigar, const uint32_t* cigar) {
	hts_pos_t l = 0;
	for (int k = 0; k < n_cigar; ++k)
		if (bam_cigar_type(bam_cigar_op(cigar[k])) & 2)
			l += bam_cigar_oplen(cigar[k]);
	return l;
}

hts_pos_t bam_endpos(const bam1_t* b) {
	hts_pos_t rlen = (b->core.flag & BAM_FUNMAP) ? 0 : bam_cigar2rlen(b->core.n_cigar, bam_get_cigar(b));
	if (rlen == 0) rlen = 1;
	return b->core.pos + rlen;
}

int cram_set_option(struct cram_fd* fd, enum hts_fmt_option opt, ...) { (void) fd; (void) opt; return -1; }

} // extern "C"
