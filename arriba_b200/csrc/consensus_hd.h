// consensus_hd.h -- read pileups and consensus sequences of the surviving fusions on the device.
//
// Behavioural contract: pileup_chimeric_alignments (output_fusions.cpp:25-107) and get_sequence_from_pileup (:109-240) as called by get_fusion_transcript_sequence
// (:242-318): per candidate two pileups (one per breakpoint) over ten views of its supporting reads, each reduced to a consensus string, the genomic position of
// every character and the bases beyond the breakpoint ("clipped"); plus the number of non-template bases between the fused segments (:300-318). What happens to the
// strings afterwards (junction marks, orientation, simplification, transcripts, peptide) stays host C++ (csrc/host/output.cpp), fed by this stage's output.
//
// One team (a thread block; one thread in the CPU stand-in) per job = (candidate, side):
//   1. every supporting read's CIGAR is walked once to find the 32-position tiles it touches (hash set in shared memory) and its introns (small registry);
//   2. the tiles are ranked (= slots in position order); the reads are walked again, one warp per read, lanes over the bases of a block: counters per position for the
//      eight frequent symbols A C G T N - > < as 16-bit halves of shared-memory words; everything else (inserted strings, IUPAC codes) goes to a short side list;
//   3. one thread per position: intron depth from the registry, coverage, and the consensus call of the column (positions with side-list entries are left to 4.);
//   4. one thread runs the reference's column automaton (gap / intron markers, clipped bases) over the calls and writes the strings.
// A job that does not fit (more tiles than the launch has, side lists full, counters beyond 16 bit, an empty insertion key) is flagged: the small launch's overflow is
// repeated by a launch with four times the tiles, what is left after that is computed by the host code the same way as before.
#pragma once
#include "model.h"
#include "prims.h"
#include "events_hd.h"

namespace arb {

struct team_t { // the threads that work on one job
	u32 rank, size;
	ARB_HD u32 lane() const { return rank & 31u; }
	ARB_HD u32 n_lanes() const { return size < 32u ? size : 32u; }
	ARB_HD u32 warp() const { return rank >> 5; }
	ARB_HD u32 n_warps() const { return size < 32u ? 1u : size >> 5; }
	ARB_HD void sync() const {
#ifdef __CUDA_ARCH__
		__syncthreads();
#endif
	}
};

enum { CJ_OK = 0, CJ_NEEDS_MORE_TILES = 1, CJ_HOST = 2, // verdict of a job; with CJ_HOST the reason:
       CJ_SIDE_LIST_FULL = 4, CJ_INTRONS_FULL = 8, CJ_EMPTY_KEY = 16, CJ_COUNTER_RANGE = 32, CJ_OUTPUT_FULL = 64, CJ_ODD_COLUMN = 128 };
enum { CS_N_TILES = 0, CS_N_OVF = 1, CS_FLAGS = 2, CS_PEAK = 3, CS_WORDS = 8 };
enum { CONS_INTRONS = 64, CONS_OVF = 96, CONS_TILE_EMPTY = 0x7FFFFFFF };
// frequent symbols, in the order of their ASCII codes where that matters: index -> character
enum { SY_A = 0, SY_C = 1, SY_G = 2, SY_T = 3, SY_N = 4, SY_DEL = 5, SY_OPEN = 6 /* > */, SY_CLOSE = 7 /* < */ };

struct ovf_entry { i32 pos; u32 a; u16 off; u8 len; u8 rc; }; // a key that is not one of the frequent symbols: bases [off, off + len) of alignment a's sequence

struct pile_space { // shared memory of a team
	u32 tiles_cap, hash_mask, P;
	i32* tile_key; u16* tile_rank; i32* tile_sorted; u32* cnt; u16* total; u8* ch; u32* ovf_mask; u16* gap_depth;
	unsigned long long* intron_key; u32* intron_cnt; ovf_entry* ovf; u32* ctl;
	static ARB_HD size_t bytes(u32 tiles) { const size_t P = (size_t) tiles * 32, H = (size_t) tiles * 2; return H * 4 + tiles * 4 + P * 16 + tiles * 4 + CONS_INTRONS * 8 + CONS_INTRONS * 4 + CONS_OVF * sizeof(ovf_entry) + CS_WORDS * 4 + P * 2 + H * 2 + tiles * 2 + P + 64; }
	ARB_HD void carve(void* base, u32 tiles) {
		tiles_cap = tiles; hash_mask = tiles * 2 - 1; P = tiles * 32;
		char* p = (char*) base;
		intron_key = (unsigned long long*) p; p += CONS_INTRONS * 8;
		tile_key = (i32*) p; p += (size_t) tiles * 2 * 4;
		tile_sorted = (i32*) p; p += (size_t) tiles * 4;
		cnt = (u32*) p; p += (size_t) P * 16;
		ovf_mask = (u32*) p; p += (size_t) tiles * 4;
		intron_cnt = (u32*) p; p += CONS_INTRONS * 4;
		ovf = (ovf_entry*) p; p += CONS_OVF * sizeof(ovf_entry);
		ctl = (u32*) p; p += CS_WORDS * 4;
		total = (u16*) p; p += (size_t) P * 2;
		tile_rank = (u16*) p; p += (size_t) tiles * 2 * 2;
		gap_depth = (u16*) p; p += (size_t) tiles * 2;
		ch = (u8*) p;
	}
};

struct consensus_job_out { u32 seq_len, pos_len, clip_len, verdict, region, capacity, launch; }; // region: where the launch that decided the job put its strings

struct consensus_stage {
	cand_state c; frag_view f; annot_view an;
	const u32* l1o; const u32* l1; const u32* l2o; const u32* l2; const u32* ldo; const u32* ld;
	const u32* rows;          // candidate of row r; job j = (rows[j >> 1], side j & 1)
	const u32* job_list;      // jobs of this launch (NULL: all of them)
	const u32* region;        // exclusive scan of the capacities U of the launch's jobs: seq at chars[2 region], clipped at chars[2 region + U], positions at pos[region]
	char* chars; i32* pos; consensus_job_out* out; u32 launch;

	// ---- the ten views of a candidate's reads (output_fusions.cpp:262-271)
	struct source { const u32* list; u32 lo, hi; u32 mate; bool rc; };
	ARB_HD void sources_of(u32 k, u32 side, source s[5]) const {
		const bool same = c.dir1[k] == c.dir2[k];
		const source A = {l1, l1o[k], l1o[k + 1], 0, false}, B = {l2, l2o[k], l2o[k + 1], 0, false}, D = {ld, ldo[k], ldo[k + 1], 0, false};
		if (side == 0) { s[0] = A; s[0].mate = SPLIT_READ; s[1] = A; s[1].mate = MATE1; s[2] = B; s[2].mate = SUPPLEMENTARY; s[2].rc = same; }
		else { s[0] = A; s[0].mate = SUPPLEMENTARY; s[0].rc = same; s[1] = B; s[1].mate = SPLIT_READ; s[2] = B; s[2].mate = MATE1; }
		s[3] = D; s[3].mate = MATE1; s[4] = D; s[4].mate = MATE2;
	}
	ARB_HD u32 capacity(u32 k, u32 tiles) const { // characters a consensus of this candidate can have, at most
		const u64 reads = (u64) (l1o[k + 1] - l1o[k]) * 2 + (u64) (l2o[k + 1] - l2o[k]) * 2 + (u64) (ldo[k + 1] - ldo[k]) * 2;
		const u64 columns = hd_min<u64>(reads * 1024, (u64) tiles * 33); // a read rarely covers more than a few hundred positions; the writer checks anyway
		return (u32) (2 * columns + 64);
	}
	ARB_HD bool takes_part(u32 frag, u32 mate, u32 direction, i32 bp) const { // output_fusions.cpp:31-52
		if (f.filter[frag] == F_duplicates) return false;
		const u32 a = f.idx(frag, mate); const bool fwd = f.fwd(a);
		if (f.n_aln[frag] == 2 && !((direction == DOWNSTREAM && fwd && f.end[a] <= bp + 2 && f.end[a] >= bp - 200) || (direction == UPSTREAM && !fwd && f.start[a] >= bp - 2 && f.start[a] <= bp + 200))) return false;
		if (f.n_aln[frag] == 3 && (mate == SPLIT_READ || mate == SUPPLEMENTARY) && f.start[a] != bp && f.end[a] != bp) return false;
		return true;
	}

	// ---- CIGAR walk of one read (output_fusions.cpp:54-98); the sink sees runs, not bases. All lanes of a warp walk the same read.
	template <class S> ARB_HD void walk(u32 frag, u32 mate, S& sink) const {
		const u32 a = f.idx(frag, mate); const bool fwd = f.fwd(a);
		const u32 sa = f.idx(frag, mate == SUPPLEMENTARY ? SPLIT_READ : mate);
		const i32 seq_size = (i32) f.seq_len[sa];
		i32 read_off = 0, ref_off = f.start[a]; i32 carry = 0; // carry: one base was already consumed by a preceding insertion
		const u32* cg = f.cig(a); const u32 nc = f.cigar_cnt[a]; const bool three = f.n_aln[frag] == 3;
		for (u32 k = 0; k < nc; ++k) {
			const u32 op = cig_op(cg[k]); const i32 len = (i32) cig_len(cg[k]);
			bool as_match = false;
			switch (op) {
				case C_I: sink.insertion(ref_off, sa, read_off, len + 1, seq_size); read_off += len + 1; ++ref_off; carry = 1; break;
				case C_N: { const i32 s0 = ref_off; ref_off += len - carry; sink.intron(s0, ref_off - 1); carry = 0; break; }
				case C_D: { const i32 n = len - carry; if (n > 0) { sink.deletion(ref_off, n); ref_off += n; } carry = 0; break; }
				case C_H: if (mate == SUPPLEMENTARY) read_off += len; break;
				case C_S:
					if (three && mate == SPLIT_READ && ((k == 0 && fwd) || (k == nc - 1 && !fwd))) { if (k == 0 && fwd) ref_off -= len; as_match = true; } // the clipped segment joins the pileup
					else read_off += len - carry;
					break;
				case C_M: case C_EQ: case C_X: as_match = true; break;
				default: break;
			}
			if (as_match) {
				const i32 run = len - carry;
				if (run > 0) {
					if (read_off < 0) sink.flag(CJ_HOST | CJ_EMPTY_KEY);
					else { // positions beyond the stored sequence enter the reference's pileup under an empty key (an insertion followed by an intron loses its carry)
						const i32 inside = hd_max(0, hd_min(run, seq_size - read_off));
						if (inside > 0) sink.match(ref_off, inside, sa, read_off, seq_size);
						if (inside < run) sink.empty(ref_off + inside, run - inside);
					}
					read_off += run; ref_off += run;
				}
				carry = 0;
			}
		}
	}

	// ---- shared-memory tables
	static ARB_HD u32 tile_hash(i32 tile) { return (u32) tile * 2654435761u >> 7; }
	static ARB_HD void touch(const pile_space& s, i32 tile) {
		u32 h = tile_hash(tile) & s.hash_mask;
		for (u32 probes = 0; probes <= s.hash_mask; ++probes, h = (h + 1) & s.hash_mask) {
			const u32 seen = ((volatile u32*) s.tile_key)[h];
			if (seen == (u32) tile) return;
			if (seen != (u32) CONS_TILE_EMPTY) continue;
			if (((volatile u32*) s.ctl)[CS_N_TILES] >= s.tiles_cap) break;
			const u32 old = atomic_cas_u32((u32*) &s.tile_key[h], (u32) CONS_TILE_EMPTY, (u32) tile);
			if (old == (u32) CONS_TILE_EMPTY) { if (atomic_add_u32(&s.ctl[CS_N_TILES], 1) >= s.tiles_cap) atomic_or_u32(&s.ctl[CS_FLAGS], CJ_NEEDS_MORE_TILES); return; }
			if (old == (u32) tile) return;
		}
		atomic_or_u32(&s.ctl[CS_FLAGS], CJ_NEEDS_MORE_TILES);
	}
	static ARB_HD u32 slot_of(const pile_space& s, i32 pos) { // index of a position whose tile was touched in pass 1
		const i32 tile = pos >> 5;
		u32 h = tile_hash(tile) & s.hash_mask;
		for (u32 probes = 0; s.tile_key[h] != tile; ++probes, h = (h + 1) & s.hash_mask) if (probes > s.hash_mask) { atomic_or_u32(&s.ctl[CS_FLAGS], CJ_HOST | CJ_ODD_COLUMN); return 0; } // cannot happen: pass 1 touched it
		return (u32) s.tile_rank[h] * 32 + ((u32) pos & 31u);
	}
	static ARB_HD void count(const pile_space& s, u32 slot, u32 symbol, u32 n = 1) { atomic_add_u32(&s.cnt[(symbol >> 1) * s.P + slot], n << ((symbol & 1) * 16)); }
	static ARB_HD u32 counted(const pile_space& s, u32 slot, u32 symbol) { return (s.cnt[(symbol >> 1) * s.P + slot] >> ((symbol & 1) * 16)) & 0xFFFFu; }
	static ARB_HD void add_overflow(const pile_space& s, i32 pos, u32 a, i32 off, i32 len, bool rc) {
		if (off > 65535 || len > 255) { atomic_or_u32(&s.ctl[CS_FLAGS], CJ_HOST | CJ_SIDE_LIST_FULL); return; }
		const u32 x = atomic_add_u32(&s.ctl[CS_N_OVF], 1);
		if (x >= CONS_OVF) { atomic_or_u32(&s.ctl[CS_FLAGS], CJ_HOST | CJ_SIDE_LIST_FULL); return; }
		ovf_entry e; e.pos = pos; e.a = a; e.off = (u16) off; e.len = (u8) len; e.rc = rc ? 1 : 0;
		s.ovf[x] = e;
		const u32 slot = slot_of(s, pos);
		atomic_or_u32(&s.ovf_mask[slot >> 5], 1u << (slot & 31u));
	}
	static ARB_HD void add_intron(const pile_space& s, i32 s0, i32 e) {
		const unsigned long long key = (unsigned long long) (u32) s0 << 32 | (u32) e;
		u32 h = ((u32) s0 * 2654435761u ^ (u32) e * 40503u) & (CONS_INTRONS - 1);
		for (u32 probes = 0; probes < CONS_INTRONS; ++probes, h = (h + 1) & (CONS_INTRONS - 1)) {
			const unsigned long long old = atomic_cas_u64((u64*) &s.intron_key[h], ~0ull, key);
			if (old == ~0ull || old == key) { atomic_add_u32(&s.intron_cnt[h], 1); return; }
		}
		atomic_or_u32(&s.ctl[CS_FLAGS], CJ_HOST | CJ_INTRONS_FULL);
	}
	ARB_HD u32 base_code(u32 a, i32 seq_size, i32 off, bool rc) const { const u8* q = f.sq(a); return rc ? nt16_complement(nt16_at(q, (u32) (seq_size - 1 - off))) : nt16_at(q, (u32) off); }
	static ARB_HD int symbol_of_code(u32 code) { switch (code) { case NT_A: return SY_A; case NT_C: return SY_C; case NT_G: return SY_G; case NT_T: return SY_T; case NT_N: return SY_N; default: return -1; } }

	struct touch_sink { // pass 1: tiles and introns
		const pile_space& s; team_t t;
		ARB_HD void span(i32 ref, i32 n) { for (i32 tile = (ref >> 5) + (i32) t.lane(); tile <= ((ref + n - 1) >> 5); tile += (i32) t.n_lanes()) touch(s, tile); }
		ARB_HD void match(i32 ref, i32 run, u32, i32, i32) { span(ref, run); }
		ARB_HD void deletion(i32 ref, i32 n) { span(ref, n); }
		ARB_HD void empty(i32 ref, i32 n) { span(ref, n); }
		ARB_HD void intron(i32 s0, i32 e) { if (t.lane() == 0) { touch(s, s0 >> 5); touch(s, e >> 5); add_intron(s, s0, e); } }
		ARB_HD void insertion(i32 pos, u32, i32, i32, i32) { if (t.lane() == 0) touch(s, pos >> 5); }
		ARB_HD void flag(u32 v) { if (t.lane() == 0) atomic_or_u32(&s.ctl[CS_FLAGS], v); }
	};
	struct count_sink { // pass 2: counters
		const consensus_stage& st; const pile_space& s; team_t t; bool rc;
		ARB_HD void base(i32 pos, u32 a, i32 seq_size, i32 off) {
			const u32 code = st.base_code(a, seq_size, off, rc); const int sy = symbol_of_code(code);
			if (sy >= 0) count(s, slot_of(s, pos), (u32) sy); else add_overflow(s, pos, a, off, 1, rc);
		}
		ARB_HD void match(i32 ref, i32 run, u32 a, i32 off, i32 seq_size) { for (i32 b = (i32) t.lane(); b < run; b += (i32) t.n_lanes()) base(ref + b, a, seq_size, off + b); }
		ARB_HD void deletion(i32 ref, i32 n) { for (i32 b = (i32) t.lane(); b < n; b += (i32) t.n_lanes()) count(s, slot_of(s, ref + b), SY_DEL); }
		ARB_HD void empty(i32 ref, i32 n) { if (t.lane() == 0) for (i32 b = 0; b < n; ++b) add_overflow(s, ref + b, 0, 0, 0, false); }
		ARB_HD void intron(i32, i32) {}
		ARB_HD void insertion(i32 pos, u32 a, i32 off, i32 n, i32 seq_size) { // key: the inserted bases and the one after them, as far as the read goes (output_fusions.cpp:61, piece())
			if (t.lane() != 0) return;
			const i32 len = off <= seq_size ? hd_min(n, seq_size - off) : 0;
			if (len <= 0) add_overflow(s, pos, 0, 0, 0, false);
			else if (len == 1) base(pos, a, seq_size, off);
			else add_overflow(s, pos, a, off, len, rc);
		}
		ARB_HD void flag(u32 v) { if (t.lane() == 0) atomic_or_u32(&s.ctl[CS_FLAGS], v); }
	};

	// ---- a column: keys in ascending order with their counts, the way the reference iterates its map<string, count>
	ARB_HD char key_char(const ovf_entry& e, u32 i) const { return nt16_char(base_code(e.a, (i32) f.seq_len[e.a], (i32) e.off + (i32) i, e.rc != 0)); }
	ARB_HD char first_char(const ovf_entry& e) const { return e.len ? key_char(e, 0) : (char) 0; }
	ARB_HD int compare_keys(const ovf_entry& x, const ovf_entry& y) const {
		const u32 n = hd_min<u32>(x.len, y.len);
		for (u32 i = 0; i < n; ++i) { const unsigned char cx = (unsigned char) key_char(x, i), cy = (unsigned char) key_char(y, i); if (cx != cy) return cx < cy ? -1 : 1; }
		return x.len < y.len ? -1 : x.len > y.len ? 1 : 0;
	}
	static ARB_HD char lower(char ch) { return ch >= 'A' && ch <= 'Z' ? (char) (ch + 32) : ch; }
	static ARB_HD char upper(char ch) { return ch >= 'a' && ch <= 'z' ? (char) (ch - 32) : ch; }
	static ARB_HD bool intron_symbol(char k) { return k == '_' || k == '>' || k == '<'; }
	struct column_entry { i32 key; u32 n; }; // key >= 0: a one-character key; key < 0: -1 - index of a side-list entry with that (longer) key
	// does `entry` take over from `best` (output_fusions.cpp:150-160)? one-character keys only on both sides of the tie rules, longer keys win by count alone
	static ARB_HD bool takes_over(bool have, char k, bool k_single, u32 n, char bk, bool b_single, u32 bn, char ref_base) {
		if (!have || n > bn) return true;
		if (n != bn) return false;
		const bool b_intron = b_single && intron_symbol(bk);
		if (k_single && k == ref_base && !b_intron) return true;
		if (k_single && k == '<' && !(b_single && (bk == '_' || bk == '>'))) return true;
		if (k_single && (k == '_' || k == '>')) return true;
		return false;
	}
	ARB_HD char ref_base_at(u32 contig, i32 pos) const { return (an.contig_seq_off[contig] != ~(u64) 0 && (u32) pos < an.contig_len[contig]) ? an.assembly[an.contig_seq_off[contig] + (u32) pos] : 'N'; }
	// the call of a column made of frequent symbols only: the character the consensus prints ('_' '>' '<' for intron columns), before the clipped / sequence decision
	ARB_HD char quick_call(const u32 n[9] /* - < > A C G N T _ */, char ref_base) const {
		const char keys[9] = {'-', '<', '>', 'A', 'C', 'G', 'N', 'T', '_'};
		bool have = false; char bk = 0; u32 bn = 0, cov = 0;
		for (int x = 0; x < 9; ++x) {
			if (n[x] == 0) continue;
			if (takes_over(have, keys[x], true, n[x], bk, true, bn, ref_base)) { have = true; bk = keys[x]; bn = n[x]; }
			if (!intron_symbol(keys[x])) cov += n[x];
		}
		const bool keep = (intron_symbol(bk) && bn >= cov) || 4ull * bn >= 3ull * cov || bk == ref_base;
		char call = keep ? bk : '?';
		if (!intron_symbol(call) && call != ref_base && ref_base != 'N') call = lower(call);
		return call;
	}

	// ---- the output of one job
	struct emitter {
		char* seq; i32* pos; char* clip; u32 cap, n_seq, n_pos, n_clip; bool overflow;
		ARB_HD void put_seq(char ch) { if (n_seq < cap) seq[n_seq] = ch; else overflow = true; ++n_seq; }
		ARB_HD void put_pos(i32 p) { if (n_pos < cap) pos[n_pos] = p; else overflow = true; ++n_pos; }
		ARB_HD void put_clip(char ch) { if (n_clip < cap) clip[n_clip] = ch; else overflow = true; ++n_clip; }
		ARB_HD void marker(const char* text, u32 n) { for (u32 i = 0; i < n; ++i) { put_seq(text[i]); put_pos(-1); } }
	};

	// the column at `slot` has side-list entries: merge them with the counters in key order and make the call (thread 0 only)
	ARB_HD void slow_column(const pile_space& s, u32 slot, i32 position, u32 depth, char ref_base, bool clipped, emitter& out, u32& verdict) const {
		column_entry col[9 + 8]; u32 n_col = 0;
		const char keys[9] = {'-', '<', '>', 'A', 'C', 'G', 'N', 'T', '_'};
		const u32 cnts[9] = {counted(s, slot, SY_DEL), counted(s, slot, SY_CLOSE), counted(s, slot, SY_OPEN), counted(s, slot, SY_A), counted(s, slot, SY_C), counted(s, slot, SY_G), counted(s, slot, SY_N), counted(s, slot, SY_T), depth};
		for (int x = 0; x < 9; ++x) if (cnts[x]) { col[n_col].key = keys[x]; col[n_col].n = cnts[x]; ++n_col; }
		const u32 n_ovf = hd_min<u32>(s.ctl[CS_N_OVF], CONS_OVF);
		for (u32 x = 0; x < n_ovf; ++x) {
			const ovf_entry& e = s.ovf[x];
			if (e.pos != position) continue;
			// find the place of this key: first entry that is not smaller
			u32 at = 0; int cmp = 1;
			for (; at < n_col; ++at) {
				if (col[at].key >= 0) { // one character c against the key: c < key iff c <= key[0] (c itself is never a side-list key); the empty key is the smallest
					const unsigned char c0 = (unsigned char) col[at].key, k0 = e.len ? (unsigned char) key_char(e, 0) : 0;
					cmp = e.len == 0 ? 1 : (e.len == 1 && c0 == k0) ? 0 : (c0 <= k0 ? -1 : 1);
				} else cmp = compare_keys(s.ovf[-1 - col[at].key], e);
				if (cmp >= 0) break;
			}
			if (at < n_col && cmp == 0) { ++col[at].n; continue; }
			if (n_col >= 9 + 8) { verdict |= CJ_HOST | CJ_ODD_COLUMN; return; }
			for (u32 y = n_col; y > at; --y) col[y] = col[y - 1];
			col[at].key = -1 - (i32) x; col[at].n = 1; ++n_col;
		}
		bool have = false; u32 best = 0, cov = 0;
		for (u32 x = 0; x < n_col; ++x) {
			const bool single = col[x].key >= 0 || s.ovf[-1 - col[x].key].len == 1;
			const char k = col[x].key >= 0 ? (char) col[x].key : first_char(s.ovf[-1 - col[x].key]);
			const bool b_single = have && (col[best].key >= 0 || s.ovf[-1 - col[best].key].len == 1);
			const char bk = !have ? 0 : col[best].key >= 0 ? (char) col[best].key : first_char(s.ovf[-1 - col[best].key]);
			if (takes_over(have, k, single, col[x].n, bk, b_single, have ? col[best].n : 0, ref_base)) { have = true; best = x; }
			if (!(single && intron_symbol(k))) cov += col[x].n;
		}
		const bool b_multi = col[best].key < 0 && s.ovf[-1 - col[best].key].len > 1;
		const bool b_empty = col[best].key < 0 && s.ovf[-1 - col[best].key].len == 0;
		const char bk = col[best].key >= 0 ? (char) col[best].key : first_char(s.ovf[-1 - col[best].key]);
		const u32 bn = col[best].n;
		const bool b_one = col[best].key >= 0 || s.ovf[-1 - col[best].key].len == 1;
		const bool keep = (b_one && intron_symbol(bk) && bn >= cov) || 4ull * bn >= 3ull * cov || (b_one && bk == ref_base);
		if (keep && b_empty) { if (!clipped) out.put_pos(position); return; } // the call is the empty string: a position without a character, like the reference's
		if (!keep || !b_multi) { // a one-character call after all
			char call = keep ? bk : '?';
			if (intron_symbol(call)) { verdict |= CJ_HOST | CJ_ODD_COLUMN; return; } // an intron call in a column with side-list keys: leave the automaton's bookkeeping to the host
			if (call != ref_base && ref_base != 'N') call = lower(call);
			if (clipped) out.put_clip(call); else { out.put_seq(call); out.put_pos(position); }
			return;
		}
		// an insertion: [inserted]next, lower case; the base after the insertion is upper case when it matches the reference (output_fusions.cpp:204-212)
		const ovf_entry& e = s.ovf[-1 - col[best].key];
		const u32 n = e.len;
		for (u32 i = 0; i < n + 1; ++i) out.put_pos(-1); // the reference grows the positions even when the call goes to the clipped bases
		for (u32 i = 0; i < n + 2; ++i) {
			char ch;
			if (i == 0) ch = '['; else if (i == n) ch = ']';
			else if (i < n) ch = lower(key_char(e, i - 1));
			else { ch = lower(key_char(e, n - 1)); if (upper(ch) == ref_base) ch = upper(ch); }
			if (clipped) out.put_clip(ch); else out.put_seq(ch);
		}
		if (!clipped) out.put_pos(position);
	}

	ARB_HD u32 intron_depth(const pile_space& s, i32 p) const {
		u32 d = 0;
		for (u32 h = 0; h < CONS_INTRONS; ++h) { const unsigned long long k = s.intron_key[h]; if (k != ~0ull && (i32) (u32) (k >> 32) < p && p < (i32) (u32) k) d += s.intron_cnt[h]; }
		return d;
	}

	// ---- one job
	ARB_HD void run(u32 j, const team_t& t, const pile_space& s) const {
		const u32 job = job_list ? job_list[j] : j;
		const u32 k = rows[job >> 1], side = job & 1;
		const u32 direction = side == 0 ? c.dir1[k] : c.dir2[k]; const i32 bp = side == 0 ? c.bp1[k] : c.bp2[k];
		const u32 contig = an.gene_contig[side == 0 ? c.gene1[k] : c.gene2[k]];
		source src[5]; sources_of(k, side, src);
		u32 first_read[6]; first_read[0] = 0; for (int x = 0; x < 5; ++x) first_read[x + 1] = first_read[x] + (src[x].hi - src[x].lo);
		const u32 n_reads = first_read[5];
		// clean tables
		for (u32 x = t.rank; x <= s.hash_mask; x += t.size) s.tile_key[x] = CONS_TILE_EMPTY;
		for (u32 x = t.rank; x < CONS_INTRONS; x += t.size) { s.intron_key[x] = ~0ull; s.intron_cnt[x] = 0; }
		for (u32 x = t.rank; x < CS_WORDS; x += t.size) s.ctl[x] = 0;
		t.sync();
		// pass 1
		{
			touch_sink sink = {s, t};
			for (u32 r = t.warp(); r < n_reads; r += t.n_warps()) {
				int x = 0; while (r >= first_read[x + 1]) ++x;
				const u32 frag = src[x].list[src[x].lo + (r - first_read[x])];
				if (takes_part(frag, src[x].mate, direction, bp)) walk(frag, src[x].mate, sink);
			}
		}
		t.sync();
		u32 verdict = s.ctl[CS_FLAGS];
		const u32 n_tiles = hd_min(s.ctl[CS_N_TILES], s.tiles_cap);
		if (n_reads > 60000) verdict |= CJ_HOST | CJ_COUNTER_RANGE; // the counters are 16 bits wide
		t.sync();
		if (verdict) { if (t.rank == 0) { consensus_job_out o = {0, 0, 0, (verdict & CJ_HOST) ? (verdict & ~(u32) CJ_NEEDS_MORE_TILES) : (u32) CJ_NEEDS_MORE_TILES, 0, 0, launch}; out[job] = o; } return; }
		// slots in position order
		for (u32 h = t.rank; h <= s.hash_mask; h += t.size) {
			const i32 key = s.tile_key[h];
			if (key == CONS_TILE_EMPTY) continue;
			u32 r = 0;
			for (u32 y = 0; y <= s.hash_mask; ++y) if (s.tile_key[y] < key) ++r; // CONS_TILE_EMPTY is larger than every tile
			s.tile_rank[h] = (u16) r; s.tile_sorted[r] = key;
		}
		const u32 n_slots = n_tiles * 32;
		for (u32 w = 0; w < 4; ++w) for (u32 x = t.rank; x < n_slots; x += t.size) s.cnt[w * s.P + x] = 0;
		for (u32 x = t.rank; x < n_tiles; x += t.size) s.ovf_mask[x] = 0;
		t.sync();
		// pass 2
		for (u32 r = t.warp(); r < n_reads; r += t.n_warps()) {
			int x = 0; while (r >= first_read[x + 1]) ++x;
			const u32 frag = src[x].list[src[x].lo + (r - first_read[x])];
			if (!takes_part(frag, src[x].mate, direction, bp)) continue;
			count_sink sink = {*this, s, t, src[x].rc};
			walk(frag, src[x].mate, sink);
		}
		t.sync();
		for (u32 h = t.rank; h < CONS_INTRONS; h += t.size) { // intron ends (output_fusions.cpp:100-106)
			const unsigned long long key = s.intron_key[h];
			if (key == ~0ull) continue;
			if (s.intron_cnt[h] > 60000) { atomic_or_u32(&s.ctl[CS_FLAGS], CJ_HOST | CJ_COUNTER_RANGE); continue; }
			count(s, slot_of(s, (i32) (u32) (key >> 32)), SY_OPEN, s.intron_cnt[h]); count(s, slot_of(s, (i32) (u32) key), SY_CLOSE, s.intron_cnt[h]);
		}
		t.sync();
		// columns
		for (u32 x = t.rank; x < n_slots; x += t.size) {
			const i32 position = s.tile_sorted[x >> 5] * 32 + (i32) (x & 31u);
			const u32 depth = intron_depth(s, position);
			u32 n_side = 0;
			if (s.ovf_mask[x >> 5] >> (x & 31u) & 1u) { const u32 n_ovf = hd_min<u32>(s.ctl[CS_N_OVF], CONS_OVF); for (u32 y = 0; y < n_ovf; ++y) if (s.ovf[y].pos == position) ++n_side; }
			const u32 n[9] = {counted(s, x, SY_DEL), counted(s, x, SY_CLOSE), counted(s, x, SY_OPEN), counted(s, x, SY_A), counted(s, x, SY_C), counted(s, x, SY_G), counted(s, x, SY_N), counted(s, x, SY_T), depth};
			u32 tot = n_side; for (int y = 0; y < 9; ++y) tot += n[y];
			if (tot > 65535) { atomic_or_u32(&s.ctl[CS_FLAGS], CJ_HOST | CJ_COUNTER_RANGE); tot = 65535; }
			s.total[x] = (u16) tot;
			s.ch[x] = tot == 0 ? 0 : n_side ? 1 : (u8) quick_call(n, ref_base_at(contig, position));
			if (tot) {
#ifdef __CUDA_ARCH__
				atomicMax(&s.ctl[CS_PEAK], tot);
#else
				if (tot > s.ctl[CS_PEAK]) s.ctl[CS_PEAK] = tot;
#endif
			}
		}
		for (u32 x = t.rank; x < n_tiles; x += t.size) { // the stretch between two tiles lies inside introns or is empty
			u32 d = 0;
			if (x > 0 && s.tile_sorted[x] > s.tile_sorted[x - 1] + 1) d = intron_depth(s, (s.tile_sorted[x - 1] + 1) * 32);
			if (d > 65535) { atomic_or_u32(&s.ctl[CS_FLAGS], CJ_HOST | CJ_COUNTER_RANGE); d = 65535; }
			s.gap_depth[x] = (u16) d;
			if (d) {
#ifdef __CUDA_ARCH__
				atomicMax(&s.ctl[CS_PEAK], d);
#else
				if (d > s.ctl[CS_PEAK]) s.ctl[CS_PEAK] = d;
#endif
			}
		}
		t.sync();
		if (t.rank != 0) return;
		// the reference's walk over the columns (output_fusions.cpp:113-237), one thread
		verdict = s.ctl[CS_FLAGS];
		const u32 U = region[j + 1] - region[j];
		emitter o; o.seq = chars + 2 * (size_t) region[j]; o.clip = o.seq + U; o.pos = pos + region[j]; o.cap = U; o.n_seq = o.n_pos = o.n_clip = 0; o.overflow = false;
		const u32 n_columns = n_tiles * 33; // per tile: the stretch before it, then its 32 positions
		const float threshold = (float) s.ctl[CS_PEAK] * 0.10f;
#define ARB_COLUMN(cx, exists, cov, first_pos, last_pos, slot) \
		const u32 tile_ = (cx) / 33, r_ = (cx) % 33; const u32 slot = r_ ? tile_ * 32 + r_ - 1 : 0; \
		const u32 cov = r_ ? s.total[slot] : s.gap_depth[tile_]; const bool exists = cov != 0; \
		const i32 first_pos = r_ ? s.tile_sorted[tile_] * 32 + (i32) r_ - 1 : (tile_ ? (s.tile_sorted[tile_ - 1] + 1) * 32 : 0), last_pos = r_ ? first_pos : s.tile_sorted[tile_] * 32 - 1;
		u32 first = n_columns, last = n_columns; // first column of the walk, one past its last
		for (u32 cx = 0; cx < n_columns; ++cx) { ARB_COLUMN(cx, exists, cov, p0, p1, slot) (void) p0; (void) p1; (void) slot; if (exists) { first = cx; break; } }
		if (direction == DOWNSTREAM) {
			for (u32 cx = first; cx < n_columns; ++cx) { ARB_COLUMN(cx, exists, cov, p0, p1, slot) (void) p0; (void) p1; (void) slot; if (!exists) continue; if ((float) cov < threshold) first = cx; else break; }
		} else {
			bool any = false; u32 at = 0;
			for (u32 cx = first; cx < n_columns; ++cx) { ARB_COLUMN(cx, exists, cov, p0, p1, slot) (void) p0; (void) p1; (void) slot; if (exists && (float) cov > threshold) { at = cx; any = true; } }
			if (any) last = at + 1;
		}
		bool intron_open = false, intron_closed = true, have_previous = false; i32 previous = 0;
		for (u32 cx = first; cx < last; ++cx) {
			ARB_COLUMN(cx, exists, cov, p0, p1, slot)
			(void) cov;
			if (!exists) continue;
			if (have_previous && previous < p0 - 1 && !intron_open) o.marker("...", 3);
			have_previous = true; previous = p1;
			const char call = r_ ? (char) s.ch[slot] : '_';
			if (call == 1) { // side-list keys in this column
				if (!intron_closed) o.marker("...", 3);
				intron_open = false; intron_closed = true;
				slow_column(s, slot, p0, intron_depth(s, p0), ref_base_at(contig, p0), (direction == UPSTREAM && p0 < bp) || (direction == DOWNSTREAM && p0 > bp), o, verdict);
			} else if (call == '_') { if (!intron_open) { o.marker("...___", 6); intron_open = true; intron_closed = false; } }
			else if (call == '>') { if (!intron_open) { o.marker("___", 3); intron_open = true; intron_closed = false; } }
			else if (call == '<') { if (!intron_open) o.marker("...___", 6); intron_open = true; intron_closed = true; }
			else {
				if (!intron_closed) o.marker("...", 3);
				intron_open = false; intron_closed = true;
				if ((direction == UPSTREAM && p0 < bp) || (direction == DOWNSTREAM && p0 > bp)) o.put_clip(call); else { o.put_seq(call); o.put_pos(p0); }
			}
		}
#undef ARB_COLUMN
		if (o.overflow) verdict |= CJ_HOST | CJ_OUTPUT_FULL;
		consensus_job_out res = {o.n_seq, o.n_pos, o.n_clip, (verdict & CJ_HOST) ? (verdict & ~(u32) CJ_NEEDS_MORE_TILES) : (u32) CJ_OK, region[j], U, launch};
		out[job] = res;
	}
};

// ---- non-template bases between the fused segments (output_fusions.cpp:300-318): the most frequent surplus of clipped bases over the read length among the split
// reads, first to reach the top count wins; 0xFFFFFFFF = more distinct values than the small table holds (the host counts that row)
struct non_template_fn {
	frag_view f; const u32* rows; const u32* l1o; const u32* l1; const u32* l2o; const u32* l2; u32* result;
	ARB_HD void operator()(u32 r) const {
		const u32 k = rows[r];
		u32 value[16], count[16]; u32 n = 0; u32 best_value = 0, best_count = 0; bool overflow = false;
		for (int which = 0; which < 2 && !overflow; ++which) {
			const u32* list = which == 0 ? l1 : l2; const u32 lo = which == 0 ? l1o[k] : l2o[k], hi = which == 0 ? l1o[k + 1] : l2o[k + 1];
			for (u32 x = lo; x < hi; ++x) {
				const u32 s = f.idx(list[x], SPLIT_READ), u = f.idx(list[x], SUPPLEMENTARY);
				const u32 cs = f.fwd(s) ? f.preclip(s) : f.postclip(s), cu = f.fwd(u) ? f.postclip(u) : f.preclip(u);
				if (cs + cu < f.seq_len[s]) continue;
				const u32 unmapped = cs + cu - f.seq_len[s];
				u32 at = 0; while (at < n && value[at] != unmapped) ++at;
				if (at == n) { if (n == 16) { overflow = true; break; } value[n] = unmapped; count[n] = 0; ++n; }
				++count[at];
				if (unmapped == best_value) { best_count = count[at]; continue; } // ++count[x] > count[x] never holds
				if (count[at] > best_count) { best_value = unmapped; best_count = count[at]; }
			}
		}
		result[r] = overflow ? 0xFFFFFFFFu : best_value;
	}
};

struct consensus_capacity_fn { consensus_stage st; u32 tiles; u32* cap; ARB_HD void operator()(u32 j) const { const u32 job = st.job_list ? st.job_list[j] : j; cap[j] = st.capacity(st.rows[job >> 1], tiles); } };
struct consensus_lengths_fn { const consensus_job_out* out; u32* seq; u32* pos; u32* clip; u8* verdict;
	ARB_HD void operator()(u32 j) const { const bool ok = out[j].verdict == CJ_OK; seq[j] = ok ? out[j].seq_len : 0; pos[j] = ok ? out[j].pos_len : 0; clip[j] = ok ? out[j].clip_len : 0; verdict[j] = (u8) out[j].verdict; } };
struct consensus_pack_fn { // the strings of a job, from its region to their place in the packed output
	const consensus_job_out* out; const char* chars[2]; const i32* pos[2]; const u32* seq_off; const u32* pos_off; const u32* clip_off; char* seq_out; i32* pos_out; char* clip_out;
	ARB_HD void operator()(u32 j) const {
		if (out[j].verdict != CJ_OK) return;
		const u32 U = out[j].capacity; const char* s = chars[out[j].launch] + 2 * (size_t) out[j].region; const i32* p = pos[out[j].launch] + out[j].region;
		for (u32 i = 0; i < out[j].seq_len; ++i) seq_out[seq_off[j] + i] = s[i];
		for (u32 i = 0; i < out[j].clip_len; ++i) clip_out[clip_off[j] + i] = s[U + i];
		for (u32 i = 0; i < out[j].pos_len; ++i) pos_out[pos_off[j] + i] = p[i];
	}
};
struct consensus_retry_fn { const consensus_job_out* out; u32* flag; ARB_HD void operator()(u32 j) const { flag[j] = out[j].verdict == CJ_NEEDS_MORE_TILES ? 1u : 0u; } };
struct consensus_retry_gather_fn { const u32* flag_scan; u32* jobs; ARB_HD void operator()(u32 j) const { if (flag_scan[j + 1] != flag_scan[j]) jobs[flag_scan[j]] = j; } };

} // namespace arb
