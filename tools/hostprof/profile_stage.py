"""Samples the program counter of the host stages (there is no perf in the image): builds a copy of the CPU stand-in library with -g, runs one sample through
the Pipeline API under LD_PRELOAD=libpcsample.so and writes one .pcs file per stage (ingest, annotate, output) for report.py.

  gcc -O2 -shared -fPIC -o /tmp/libpcsample.so tools/hostprof/pcsample.c -ldl
  LD_PRELOAD=/tmp/libpcsample.so python tools/hostprof/profile_stage.py PREFIX OUTDIR     # PREFIX.bam / .gtf / .fa, e.g. a world written by build/synth
  python tools/hostprof/report.py OUTDIR/ingest.pcs OUTDIR/libhostsim_g.so

The stand-in runs the kernels as loops, so only the host stages' numbers mean anything; the timer counts CPU time of all threads."""
import ctypes, os, sys, time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from arriba_b200 import _build as B, lib  # noqa: E402


def build_debug_copy(outdir):
    objdir = os.path.join(outdir, "obj"); os.makedirs(objdir, exist_ok=True)
    flags = B.GXX_FLAGS + ["-g", "-fno-omit-frame-pointer", "-DARB_HOSTSIM", "-I", os.path.join(B.ROOT, "include")]
    jobs = [["g++", "-x", "c++"] + flags + ["-c", os.path.join(B.CSRC, s), "-o", os.path.join(objdir, s + ".o")] for s in B.CU_SOURCES]
    jobs += [["g++"] + flags + ["-c", os.path.join(B.CSRC, s), "-o", os.path.join(objdir, s.replace("/", "_") + ".o")] for s in B.CPP_SOURCES]
    with ThreadPoolExecutor(8) as ex:
        list(ex.map(B._run, jobs))
    so = os.path.join(outdir, "libhostsim_g.so")
    B._run(["g++", "-shared", "-o", so] + [os.path.join(objdir, s.replace("/", "_") + ".o") for s in B.CU_SOURCES + B.CPP_SOURCES] + ["-lz", "-lpthread"])
    return so


def main():
    prefix, outdir = sys.argv[1], sys.argv[2]
    os.makedirs(outdir, exist_ok=True)
    so = build_debug_copy(outdir)
    prof = ctypes.CDLL(None)   # the preloaded sampler
    if not hasattr(prof, "pcsample_reset"):
        raise SystemExit("run under LD_PRELOAD=libpcsample.so (see the module docstring)")
    p = lib.Pipeline(prefix + ".bam", prefix + ".gtf", prefix + ".fa", threads=os.cpu_count() or 1, lib_path=so,
                     output=os.path.join(outdir, "fusions.tsv"), discarded=os.path.join(outdir, "fusions.discarded.tsv"))
    p.step(lib.STEP_LOAD_REFERENCE)

    def sampled(name, f):
        prof.pcsample_reset(); t = time.time(); f()
        print("%-10s %.2f s" % (name, time.time() - t)); prof.pcsample_dump(os.path.join(outdir, name + ".pcs").encode())
    sampled("ingest", lambda: p.step(lib.STEP_INGEST))
    sampled("annotate", lambda: p.step(lib.STEP_ANNOTATE))
    for s in range(lib.STEP_ANNOTATE + 1, lib.STEP_COUNT):
        p.step(s)
    sampled("events", lambda: p.events(len(lib.EV_NAMES) - 1))
    sampled("output", p.write_output)
    p.close()


if __name__ == "__main__":
    main()
