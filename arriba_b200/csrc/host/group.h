// group.h -- the ranks of ONE sample on the GPUs of one node (one process per GPU): what they share on the host.
//
// Everything the host produces (fragment columns, labels, coverage, formatted rows) lives in host memory anyway, and all ranks run on the same node, so the
// host-side exchange is shared memory: a rank writes its part of a table in place and the others read it -- no serialisation, no copy, and on the way to a
// device the data crosses PCIe exactly once (host -> owner's GPU), where a route through the producer's GPU would add a second PCIe crossing and an NVLink hop.
// Device-resident tables (the candidate tables after find_fusions, the per-fragment re-alignment verdicts) are exchanged device-to-device by NCCL (comm.h).
//
// A group is a control block (rendezvous file in /dev/shm: barrier, mailboxes, the NCCL id) plus named segments (memfd_create by rank 0, opened by the peers
// through /proc/<pid>/fd, so their size is not bounded by the size of /dev/shm). Segments are recycled from sample to sample inside a process group: fresh
// pages cost ~0.25 s per GB to fault in.
#pragma once
#include <atomic>
#include <map>
#include <string>
#include <vector>
#include "../hd.h"

namespace arb { namespace host {

struct host_group {
	enum { MAX_RANKS = 16, MAIL_WORDS = 64, MAX_SEGMENTS = 256 };
	struct segment_slot { std::atomic<u32> version; int fd; u64 bytes; char tag[48]; };
	struct control {
		std::atomic<u32> magic, attached, arrived, generation, failed, nccl_id_ready;
		int creator_pid;
		char error[256];
		unsigned char nccl_id[128];
		u64 mail[MAX_RANKS][MAIL_WORDS];
		segment_slot segments[MAX_SEGMENTS]; std::atomic<u32> n_segments;
	};
	struct mapping { void* p; u64 bytes; u32 version; int slot; };

	int rank, world; std::string name; control* ctl; int ctl_fd; double timeout_s;
	std::map<std::string, mapping> maps;

	// rank 0 creates the control block, the others wait for it; `name` must be unique per job (the launcher hands it out)
	host_group(const std::string& name, int rank, int world);
	~host_group();
	void barrier();                                   // throws when a peer has called fail() or does not arrive within the time limit
	void fail(const std::string& message);            // wakes the peers out of their barriers with an error instead of leaving them to hang
	void allgather(const u64* mine, int words, u64* all /* world * words */);   // contains two barriers
	u64 sum(u64 mine); u64 exclusive_sum(u64 mine, u64* total = NULL);
	// collective: one block of `bytes` (the same value on every rank) shared by all ranks under `tag`; kept and reused by later calls that fit. Contents are
	// undefined after the call (a reused block holds the previous sample's data).
	void* segment(const std::string& tag, u64 bytes);
	template <class T> T* array(const std::string& tag, u64 n) { return (T*) segment(tag, (n ? n : 1) * sizeof(T)); }
};

}} // namespace
