// LD_PRELOAD sampling profiler: ITIMER_PROF -> SIGPROF on whichever thread is running; records PCs; dumps with module offsets at exit.
#define _GNU_SOURCE
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <ucontext.h>
#include <dlfcn.h>
#include <unistd.h>
#define MAXS (1<<22)
static void* volatile samples[MAXS]; static volatile long n_samples;
static volatile int enabled = 1;
static void on_prof(int sig, siginfo_t* si, void* uc_) { ucontext_t* uc = (ucontext_t*) uc_; if (!enabled) return; long i = __sync_fetch_and_add(&n_samples, 1); if (i < MAXS) samples[i] = (void*) uc->uc_mcontext.gregs[REG_RIP]; }
__attribute__((constructor)) static void init(void) {
	struct sigaction sa; memset(&sa, 0, sizeof sa); sa.sa_sigaction = on_prof; sa.sa_flags = SA_SIGINFO | SA_RESTART; sigaction(SIGPROF, &sa, NULL);
	struct itimerval it; it.it_interval.tv_sec = 0; it.it_interval.tv_usec = 1000; it.it_value = it.it_interval; setitimer(ITIMER_PROF, &it, NULL);
}
void pcsample_reset(void) { n_samples = 0; }
void pcsample_dump(const char* path) {
	enabled = 0;
	FILE* f = fopen(path, "w"); if (!f) return;
	long n = n_samples < MAXS ? n_samples : MAXS;
	for (long i = 0; i < n; ++i) { Dl_info di; if (dladdr(samples[i], &di) && di.dli_fname) fprintf(f, "%s\t%lx\t%s\n", di.dli_fname, (unsigned long) ((char*) samples[i] - (char*) di.dli_fbase), di.dli_sname ? di.dli_sname : "?"); else fprintf(f, "?\t%lx\t?\n", (unsigned long) samples[i]); }
	fclose(f); enabled = 1;
}
