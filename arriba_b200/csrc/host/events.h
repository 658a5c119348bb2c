// events.h -- host image of the candidate table ("fusions") and the event-level stages that stay on the host.
//
// The reference keeps candidates in an unordered_map whose ITERATION ORDER several stages depend on
// (estimate_expected_fusions' first-come partner sets, select_best's tie rules, recover_isoforms' last-wins map,
// the unsorted discarded file; SURVEY.md section 7 "hard parts" 2b). The device numbers candidates by first insertion;
// replay_iteration_order() inserts the same keys, in that order, into a libstdc++ unordered_map with the reference's
// hash (common.hpp:295-314) and reads the iteration order back, so every host loop below visits candidates exactly
// like the reference does.
#pragma once
#include <string>
#include <vector>
#include "refdata.h"
#include "ingest.h"

namespace arb { namespace host {

struct event_table {
	u32 n;
	// column<T>: resize() does not touch new elements (they are all written by the device-to-host copies that follow)
	column<u32> gene1, gene2, split_reads1, split_reads2, discordant_mates;
	column<u16> contig1, contig2;
	column<i32> bp1, bp2, anchor1, anchor2;
	column<u8> dir1, dir2, filter, bits, bits2, confidence;
	column<float> evalue;
	column<u32> list1_off, list2_off, listd_off, list1, list2, listd; // CSR, fragment indices in name order
	std::vector<u32> order; // order[k] = candidate visited k-th by the reference's loops
	std::vector<u32> rank_of; // inverse of `order`
	event_table(): n(0) {}

	bool exonic1(u32 k) const { return bits[k] & CB_EXONIC1; }
	bool exonic2(u32 k) const { return bits[k] & CB_EXONIC2; }
	bool spliced1(u32 k) const { return bits[k] & CB_SPLICED1; }
	bool spliced2(u32 k) const { return bits[k] & CB_SPLICED2; }
	u32 supporting_reads(u32 k) const { return split_reads1[k] + split_reads2[k] + discordant_mates[k]; }
	u32 n_list1(u32 k) const { return list1_off[k + 1] - list1_off[k]; }
	u32 n_list2(u32 k) const { return list2_off[k + 1] - list2_off[k]; }
	u32 n_listd(u32 k) const { return listd_off[k + 1] - listd_off[k]; }
	bool is_read_through(u32 k) const { return contig1[k] == contig2[k] && bp2[k] - bp1[k] < 400000 && dir1[k] == DOWNSTREAM && dir2[k] == UPSTREAM; } // common.hpp:265-269
};

}} // namespace
