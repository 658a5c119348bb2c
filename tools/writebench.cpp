// writebench.cpp -- how fast does a gigabyte of formatted rows reach the page cache on this box? (the writer's "file written" lap)
//   g++ -O2 -pthread -o build/writebench tools/writebench.cpp && build/writebench DIR [MiB] [threads]
// Methods: shared mapping filled by T threads (what csrc/host/output.cpp does), pwrite() by T threads in 8 MiB pieces, one write() loop, pwrite after fallocate.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/vfs.h>
#include <thread>
#include <unistd.h>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
	if (argc < 2) { fprintf(stderr, "usage: writebench DIR [MiB] [threads]\n"); return 1; }
	const std::string dir = argv[1]; const size_t bytes = (size_t) (argc > 2 ? atol(argv[2]) : 760) << 20; const int T = argc > 3 ? atoi(argv[3]) : 16;
	struct statfs fs; if (statfs(dir.c_str(), &fs) == 0) printf("filesystem type 0x%lx, block size %ld\n", (unsigned long) fs.f_type, (long) fs.f_bsize);
	std::vector<char> src(bytes); for (size_t i = 0; i < bytes; ++i) src[i] = (char) ('a' + i % 23);
	const std::string path = dir + "/writebench.tmp";
	for (int method = 0; method < 5; ++method) for (int rep = 0; rep < 2; ++rep) {
		unlink(path.c_str());
		const double t0 = now();
		const int fd = open(path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0666);
		if (fd < 0) { perror("open"); return 1; }
		const char* name = "";
		if (method == 0) {
			name = "ftruncate + shared mapping, threads memcpy";
			if (ftruncate(fd, (off_t) bytes) != 0) perror("ftruncate");
			char* map = (char*) mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
			std::vector<std::thread> pool;
			for (int t = 0; t < T; ++t) pool.emplace_back([&, t]() { const size_t lo = bytes * t / T, hi = bytes * (t + 1) / T; memcpy(map + lo, src.data() + lo, hi - lo); });
			for (auto& th : pool) th.join();
			munmap(map, bytes);
		} else if (method == 1 || method == 3) {
			name = method == 1 ? "pwrite by threads, 8 MiB pieces" : "fallocate + pwrite by threads, 8 MiB pieces";
			if (method == 3 && posix_fallocate(fd, 0, (off_t) bytes) != 0) perror("fallocate");
			std::vector<std::thread> pool;
			for (int t = 0; t < T; ++t) pool.emplace_back([&, t]() {
				const size_t lo = bytes * t / T, hi = bytes * (t + 1) / T;
				for (size_t at = lo; at < hi; ) { const size_t n = std::min<size_t>(8u << 20, hi - at); const ssize_t w = pwrite(fd, src.data() + at, n, (off_t) at); if (w <= 0) { perror("pwrite"); return; } at += (size_t) w; }
			});
			for (auto& th : pool) th.join();
		} else if (method == 2) {
			name = "one thread, write() in 8 MiB pieces";
			for (size_t at = 0; at < bytes; ) { const ssize_t w = write(fd, src.data() + at, std::min<size_t>(8u << 20, bytes - at)); if (w <= 0) { perror("write"); break; } at += (size_t) w; }
		} else {
			name = "ftruncate + shared mapping + MADV_POPULATE_WRITE by threads, then memcpy";
			if (ftruncate(fd, (off_t) bytes) != 0) perror("ftruncate");
			char* map = (char*) mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
			std::vector<std::thread> pool;
			for (int t = 0; t < T; ++t) pool.emplace_back([&, t]() {
				const size_t lo = (bytes * t / T) & ~(size_t) 4095, hi = t + 1 == T ? bytes : (bytes * (t + 1) / T) & ~(size_t) 4095;
#ifdef MADV_POPULATE_WRITE
				madvise(map + lo, hi - lo, MADV_POPULATE_WRITE);
#endif
				memcpy(map + lo, src.data() + lo, hi - lo);
			});
			for (auto& th : pool) th.join();
			munmap(map, bytes);
		}
		close(fd);
		const double t1 = now();
		printf("%-78s rep %d: %7.1f ms  (%.2f GB/s)\n", name, rep, (t1 - t0) * 1e3, bytes / (t1 - t0) * 1e-9);
	}
	unlink(path.c_str());
	return 0;
}
