#!/usr/bin/env python
"""Reference md5 sums of fusions.tsv / fusions.discarded.tsv for the bench workloads (bench.py compares the files its timed steps wrote against them).

Runs the UNMODIFIED reference (oracle/_ref/arriba, built from /root/reference by oracle/Makefile) on the synthetic world of a bench workload, on the CPU
container -- 25 to 60 minutes on one core and tens of GB of memory per workload, which is why the sums are committed instead of being made on the GPU box.
    python tests/golden/make_full_size_md5.py cfg2_10M_2x101_50k [more workloads]
ARB_DET_ALLOC=1: gene sets are ordered by pointer value in the reference; the bump arena makes that order the creation order."""
import hashlib, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from arriba_b200 import _build

OUT = os.path.join(ROOT, "tests", "golden", "full_size_md5.json")


def main():
    oracle = _build.build_oracle()
    for name in sys.argv[1:]:
        prefix = bench.ensure_world(name)
        out = prefix + ".ref_full"; os.makedirs(out, exist_ok=True)
        t0 = time.time()
        env = dict(os.environ, ARB_DET_ALLOC="1")
        r = subprocess.run([oracle, "-x", prefix + ".bam", "-g", prefix + ".gtf", "-a", prefix + ".fa", "-o", os.path.join(out, "fusions.tsv"), "-O", os.path.join(out, "fusions.discarded.tsv"), "-f", "blacklist"],
                           stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=env)
        if r.returncode != 0:
            raise SystemExit("reference failed on %s" % name)
        secs = time.time() - t0
        done = [l for l in r.stdout.splitlines() if "Done" in l or "total=" in l]
        table = json.load(open(OUT)) if os.path.exists(OUT) else {}
        entry = {"source": "oracle/_ref/arriba (unmodified reference + htslib shim), ARB_DET_ALLOC=1, -f blacklist, one core, %.0f s; %s" % (secs, " | ".join(done)[-300:])}
        for f in ("fusions.tsv", "fusions.discarded.tsv"):
            h = hashlib.md5()
            with open(os.path.join(out, f), "rb") as fh:
                for block in iter(lambda: fh.read(1 << 24), b""):
                    h.update(block)
            entry[f] = h.hexdigest(); entry[f + ".bytes"] = os.path.getsize(os.path.join(out, f))
        table[name] = entry
        json.dump(table, open(OUT, "w"), indent=1, sort_keys=True)
        print(name, entry, flush=True)


if __name__ == "__main__":
    main()
