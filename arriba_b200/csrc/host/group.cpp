// group.cpp -- see group.h.
#include "group.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/syscall.h>
#include <unistd.h>

namespace arb { namespace host {

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static const u32 GROUP_MAGIC = 0xA881BA02u;

static std::string control_path(const std::string& name) {
	std::string clean; for (size_t k = 0; k < name.size(); ++k) { const char c = name[k]; clean += (isalnum((unsigned char) c) || c == '_' || c == '-' || c == '.') ? c : '_'; }
	struct stat st;
	const char* dir = (::stat("/dev/shm", &st) == 0 && S_ISDIR(st.st_mode) && ::access("/dev/shm", W_OK) == 0) ? "/dev/shm" : "/tmp";
	return std::string(dir) + "/arb_group_" + clean;
}

host_group::host_group(const std::string& name_, int rank_, int world_): rank(rank_), world(world_), name(name_), ctl(NULL), ctl_fd(-1), timeout_s(900) {
	if (world < 1 || world > MAX_RANKS || rank < 0 || rank >= world) throw std::runtime_error("invalid rank / world size of the group");
	if (const char* s = getenv("ARB_GROUP_TIMEOUT")) timeout_s = std::max(1.0, atof(s));
	const std::string path = control_path(name);
	const double t0 = now_s();
	if (rank == 0) {
		::unlink(path.c_str());
		ctl_fd = ::open(path.c_str(), O_RDWR | O_CREAT | O_EXCL, 0600);
		if (ctl_fd < 0 || ::ftruncate(ctl_fd, sizeof(control)) != 0) throw std::runtime_error("cannot create the control block of the group at " + path);
	} else {
		for (;;) { // the creator may not have got there yet
			ctl_fd = ::open(path.c_str(), O_RDWR);
			if (ctl_fd >= 0) { struct stat st; if (::fstat(ctl_fd, &st) == 0 && (size_t) st.st_size >= sizeof(control)) break; ::close(ctl_fd); ctl_fd = -1; }
			if (now_s() - t0 > timeout_s) throw std::runtime_error("the control block of the group did not appear at " + path);
			::usleep(1000);
		}
	}
	void* p = ::mmap(NULL, sizeof(control), PROT_READ | PROT_WRITE, MAP_SHARED, ctl_fd, 0);
	if (p == MAP_FAILED) throw std::runtime_error("cannot map the control block of the group");
	ctl = (control*) p;
	if (rank == 0) {
		memset((void*) ctl, 0, sizeof(control));
		ctl->creator_pid = (int) ::getpid();
		ctl->magic.store(GROUP_MAGIC, std::memory_order_release);
	} else {
		while (ctl->magic.load(std::memory_order_acquire) != GROUP_MAGIC) { if (now_s() - t0 > timeout_s) throw std::runtime_error("the control block of the group was never initialised"); ::usleep(200); }
	}
	ctl->attached.fetch_add(1);
	while ((int) ctl->attached.load() < world) { if (now_s() - t0 > timeout_s) throw std::runtime_error("not all ranks joined the group"); ::usleep(200); }
	if (rank == 0) ::unlink(path.c_str()); // everybody holds it open: the name can go (nothing is left behind if the job dies)
}

host_group::~host_group() {
	for (std::map<std::string, mapping>::iterator m = maps.begin(); m != maps.end(); ++m) if (m->second.p) ::munmap(m->second.p, m->second.bytes);
	if (ctl) {
		if (rank == 0) for (u32 s = 0; s < ctl->n_segments.load() && s < MAX_SEGMENTS; ++s) if (ctl->segments[s].fd > 0) ::close(ctl->segments[s].fd);
		::munmap((void*) ctl, sizeof(control));
	}
	if (ctl_fd >= 0) ::close(ctl_fd);
}

void host_group::fail(const std::string& message) {
	if (!ctl) return;
	if (ctl->failed.exchange(1) == 0) { snprintf(ctl->error, sizeof(ctl->error), "rank %d: %s", rank, message.c_str()); }
}

void host_group::barrier() {
	if (world == 1) return;
	const double t0 = now_s();
	const u32 gen = ctl->generation.load(std::memory_order_acquire);
	if (ctl->arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == (u32) world) {
		ctl->arrived.store(0, std::memory_order_relaxed);
		ctl->generation.fetch_add(1, std::memory_order_release);
		return;
	}
	u32 spins = 0;
	while (ctl->generation.load(std::memory_order_acquire) == gen) {
		if (ctl->failed.load()) throw std::runtime_error(std::string("a peer of the group failed: ") + ctl->error);
		if (++spins < 200) { ::sched_yield(); continue; }
		::usleep(50);
		if ((spins & 1023) == 0 && now_s() - t0 > timeout_s) { fail("timed out in a barrier"); throw std::runtime_error("timed out waiting for the peers of the group"); }
	}
	if (ctl->failed.load()) throw std::runtime_error(std::string("a peer of the group failed: ") + ctl->error);
}

void host_group::allgather(const u64* mine, int words, u64* all) {
	if (words > MAIL_WORDS) throw std::runtime_error("mailbox too small");
	for (int k = 0; k < words; ++k) ctl->mail[rank][k] = mine[k];
	barrier();
	for (int r = 0; r < world; ++r) for (int k = 0; k < words; ++k) all[(size_t) r * words + k] = ctl->mail[r][k];
	barrier(); // nobody overwrites a mailbox before everybody has read it
}

u64 host_group::sum(u64 mine) { u64 all[MAX_RANKS]; allgather(&mine, 1, all); u64 s = 0; for (int r = 0; r < world; ++r) s += all[r]; return s; }
u64 host_group::exclusive_sum(u64 mine, u64* total) { u64 all[MAX_RANKS]; allgather(&mine, 1, all); u64 s = 0, before = 0; for (int r = 0; r < world; ++r) { if (r == rank) before = s; s += all[r]; } if (total) *total = s; return before; }

static u64 size_class(u64 bytes) { u64 c = (u64) 1 << 16; while (c < bytes) c += c / 4; return (c + 4095) & ~(u64) 4095; }

void* host_group::segment(const std::string& tag, u64 bytes) {
	// every rank sees the same sequence of (tag, bytes), so every rank takes the same decision about reuse without talking
	mapping& m = maps[tag];
	if (m.p && m.bytes >= bytes) { barrier(); return m.p; } // the barrier: nobody writes into a recycled block while a peer still reads the previous contents
	const u64 want = size_class(bytes);
	if (m.p) { ::munmap(m.p, m.bytes); m.p = NULL; }
	int slot = m.version == 0 ? -1 : m.slot;
	if (rank == 0) {
		if (slot < 0) { slot = (int) ctl->n_segments.fetch_add(1); if (slot >= MAX_SEGMENTS) { fail("too many shared segments"); throw std::runtime_error("too many shared segments"); } snprintf(ctl->segments[slot].tag, sizeof(ctl->segments[slot].tag), "%s", tag.c_str()); }
		segment_slot& s = ctl->segments[slot];
		if (s.fd > 0) ::close(s.fd);
		const int fd = (int) ::syscall(SYS_memfd_create, tag.c_str(), 0u);
		if (fd < 0 || ::ftruncate(fd, (off_t) want) != 0) { fail("cannot create a shared segment"); throw std::runtime_error("cannot create a shared segment of " + std::to_string(want) + " bytes"); }
		s.fd = fd; s.bytes = want;
		s.version.fetch_add(1, std::memory_order_release);
	}
	barrier();
	if (rank != 0 && slot < 0) { // find the slot of this tag (slots are only ever appended)
		const u32 n = ctl->n_segments.load();
		for (u32 s = 0; s < n && s < MAX_SEGMENTS; ++s) if (tag == ctl->segments[s].tag) slot = (int) s;
		if (slot < 0) { fail("shared segment not announced"); throw std::runtime_error("shared segment '" + tag + "' was not announced by rank 0"); }
	}
	segment_slot& s = ctl->segments[slot];
	int fd;
	if (rank == 0) fd = s.fd;
	else {
		char path[64]; snprintf(path, sizeof(path), "/proc/%d/fd/%d", ctl->creator_pid, s.fd);
		fd = ::open(path, O_RDWR);
		if (fd < 0) { fail("cannot open a shared segment of rank 0"); throw std::runtime_error(std::string("cannot open ") + path + " (shared segment of rank 0)"); }
	}
	void* p = ::mmap(NULL, s.bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
	if (rank != 0) ::close(fd);
	if (p == MAP_FAILED) { fail("cannot map a shared segment"); throw std::runtime_error("cannot map a shared segment"); }
	m.p = p; m.bytes = s.bytes; m.version = s.version.load(); m.slot = slot;
	barrier(); // rank 0 must not replace the descriptor before everybody has opened it
	return p;
}

}} // namespace
