#!/usr/bin/env python
"""bench.py -- chimeric fragments per second of the hot path (BAM ingest -> read-level cascade -> candidate generation ...).

One "step" = one pass of the whole path over one synthetic chimeric BAM (BASELINE.json configs[1]: 10 M fragments, 2x101 bp,
50 k breakpoints; the genome is a synthetic stand-in for hg38 at 1:10 scale because no reference genome exists offline).
  value : fragments/s of the device-resident stages, CUDA-event timed, inputs already in HBM
  e2e   : fragments/s through the public Pipeline API from the BAM file on disk (host decode, H2D, kernels, D2H of results)
  --impl reference : the unmodified reference (oracle/_ref/arriba) on the host cores, bounded sample of the same world
See DESIGN.md section "Measurement" for the definitions of the roofline numbers."""
import argparse, json, os, subprocess, sys, time, threading, hashlib, re

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: synth parameters
    "cfg2_10M_2x101_50k": dict(scale=0.1, genes=20000, breakpoints=50000, fragments=10000000, read_length=101),
    "mid_1M_2x101_5k": dict(scale=0.02, genes=4000, breakpoints=5000, fragments=1000000, read_length=101),
    "tiny_20k": dict(scale=0.001, genes=400, breakpoints=200, fragments=20000, read_length=101),
}
UNIT = "reads/s"   # chimeric fragments per second
SCOPE = "ingest..fusions.tsv"  # one step = BAM ingest -> read filters -> candidates -> event filters -> fusions.tsv + discarded.tsv (reference loading excluded)


def world_dir(name):
    base = os.environ.get("ARB_BENCH_DIR", "/tmp/arb_bench")
    return os.path.join(base, name)


def ensure_world(name, sample_breakpoints=None):
    """Generates (once) the synthetic world; returns the file prefix. A sample shares genome/GTF and keeps the per-breakpoint depth."""
    from arriba_b200 import _build
    synth = _build.build_tools()
    p = WORKLOADS[name]
    d = world_dir(name); os.makedirs(d, exist_ok=True)
    prefix = os.path.join(d, "w")
    common = [synth, "--seed", str(0xA881BA), "--scale", str(p["scale"]), "--genes", str(p["genes"]), "--breakpoints", str(p["breakpoints"]),
              "--fragments", str(p["fragments"]), "--read-length", str(p["read_length"])]
    if not os.path.exists(prefix + ".done"):
        subprocess.run(common + ["--prefix", prefix], check=True, stderr=subprocess.DEVNULL)
        open(prefix + ".done", "w").write("ok")
    if sample_breakpoints is None:
        return prefix
    sp = os.path.join(d, "sample%d" % sample_breakpoints)
    if not os.path.exists(sp + ".done"):
        subprocess.run(common + ["--prefix", sp, "--reads-only", "--emit-breakpoints", str(sample_breakpoints)], check=True, stderr=subprocess.DEVNULL)
        for ext in (".fa", ".gtf"):
            if not os.path.exists(sp + ext):
                os.symlink(prefix + ext, sp + ext)
        open(sp + ".done", "w").write("ok")
    return sp


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    def __init__(self, device):
        super().__init__(daemon=True); self.device = device; self.samples = []; self.stop_flag = False
    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + q, "--format=csv,noheader,nounits"], stdout=subprocess.PIPE, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.2)
    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None, "reasons": sorted(reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def reference_run(prefix, threads):
    """Runs the unmodified reference CLI; returns (fragments, seconds over SCOPE, total seconds)."""
    from arriba_b200 import _build
    oracle = _build.build_oracle()
    out = prefix + ".ref_out"
    os.makedirs(out, exist_ok=True)
    t0 = time.time()
    r = subprocess.run([oracle, "-x", prefix + ".bam", "-g", prefix + ".gtf", "-a", prefix + ".fa", "-o", os.path.join(out, "fusions.tsv"), "-O", os.path.join(out, "discarded.tsv"),
                        "-f", "blacklist", "-@", str(threads)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    total = time.time() - t0
    if r.returncode != 0:
        raise RuntimeError("reference run failed: " + r.stderr[-500:])
    n = int(re.search(r"\(total=(\d+)\)", r.stdout).group(1))
    # time stamps of the reference's own progress lines (1 s resolution): start of BAM reading .. first line after find_fusions
    def stamp(line):
        m = re.match(r"\[\d+-\d+-\d+T(\d+):(\d+):(\d+)\]", line)
        return int(m.group(1)) * 3600 + int(m.group(2)) * 60 + int(m.group(3))
    lines = r.stdout.splitlines()
    t_start = [stamp(l) for l in lines if "Reading chimeric alignments" in l][0]
    after = [stamp(l) for l in lines if "Freeing resources" in l]
    t_end = after[0] if after else stamp(lines[-1])
    scope_s = max(1.0, float((t_end - t_start) % 86400))
    return n, scope_s, total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2_10M_2x101_50k", choices=sorted(WORKLOADS))
    ap.add_argument("--threads", type=int, default=0, help="host threads per rank (0 = cores / ranks)")
    ap.add_argument("--sample-breakpoints", type=int, default=0, help="cpu baseline sample size in breakpoints (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sharded-extra", action="store_true", help="N>1, mode samples: skip the additional one-sample-over-all-ranks step")
    ap.add_argument("--mode", choices=["samples", "sharded"], default="samples",
                    help="N>1: 'samples' = one independent sample per GPU (weak scaling, no collective); 'sharded' = ONE sample partitioned by contig pair with two NCCL all-gathers (strong scaling)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    usable = cores   # CPUs the container may actually burn: the cgroup quota, where one is set (the GPU boxes show 128 CPUs and grant 16)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            usable = min(cores, max(1.0, float(quota) / float(period)))
    except Exception:
        pass
    threads = args.threads or max(2, min(64, int(round(2 * usable / max(1, world)))))   # 2 threads per usable CPU measured best (profiles/r01i)
    wl = WORKLOADS[args.workload]
    sample_bp = args.sample_breakpoints or max(50, int(wl["breakpoints"] * 400000 / wl["fragments"]))  # ~400 k fragments: 10-30 s of reference CPU time
    metric = "chimeric reads/sec end-to-end (ingest→fusions.tsv)"   # BASELINE.json; a "read" is one chimeric fragment (read pair + supplementary), the unit the reference counts
    config = {"workload": "synthetic %s: %d fragments 2x%d bp, %d breakpoints, genome %.0f%% of hg38 size (synthetic), %d genes" %
              (args.workload, wl["fragments"], wl["read_length"], wl["breakpoints"], wl["scale"] * 100, wl["genes"]),
              "scope": SCOPE, "value_is": "device-resident stages (CUDA events, fragment table in HBM)", "e2e_is": "BAM on disk -> both TSV files on disk through the public Pipeline API (the headline)", "l2": "inputs (>2 GB of SoA columns per step) exceed the 126 MB L2", "host_threads_per_rank": threads, "host_cpus_usable": usable, "sharding": "one independent BAM per rank (weak)"}

    if args.impl == "reference":
        if rank != 0:
            return
        prefix = ensure_world(args.workload, sample_bp)
        vals = []
        for _ in range(max(1, min(args.steps, 2))):  # one reference pass takes tens of seconds
            n, scope_s, total = reference_run(prefix, cores)
            vals.append((n / scope_s, n, scope_s, total))
        v, n, scope_s, total = max(vals)
        line = {"impl": "reference", "metric": metric, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals), "warmup": 0, "ms_per_step": scope_s * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "reference",
                                 "sample": "first %d breakpoints of the workload at full depth = %d fragments; reference's own time stamps over %s (1 s resolution); decode threads -@ %d have no effect in the shim build" % (sample_bp, n, SCOPE, cores)},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return

    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries the one JSON line only (NCCL prints its version banner where NCCL_DEBUG=VERSION)
    import torch
    from arriba_b200 import lib as L, _build
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU reference")
    if not os.path.exists(_build.PRODUCT_LIB):
        _build.build_product()
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        prefix = ensure_world(args.workload)
    if dist:
        dist.barrier()
    prefix = ensure_world(args.workload)

    def one_step(sharded=False):
        """ingest .. fusions.tsv through the public Pipeline API; returns (fragments, e2e seconds, device ms, stats, timings, d2h bytes)"""
        outdir = os.path.join(world_dir(args.workload), "out_rank%d" % rank); os.makedirs(outdir, exist_ok=True)
        p = L.Pipeline(prefix + ".bam", prefix + ".gtf", prefix + ".fa", threads=threads, device=local_rank,
                       output=os.path.join(outdir, "fusions.tsv"), discarded=os.path.join(outdir, "fusions.discarded.tsv"))
        p.step(L.STEP_LOAD_REFERENCE)          # genome + annotation: loaded once per run in a real deployment, outside the timed region
        if sharded:
            from arriba_b200 import sharded as S
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            S.run_sharded(p, rank, world, reference_loaded=True)   # two NCCL all-gathers inside; rank 0 writes the files
            e2e_s = time.perf_counter() - t0
        else:
            t0 = time.perf_counter()
            for s in range(L.STEP_INGEST, L.STEP_COUNT):
                p.step(s)
            p.events(len(L.EV_NAMES) - 1)          # event-level chain incl. the device stages and the D2H of the candidate table
            p.write_output()
            e2e_s = time.perf_counter() - t0
        ctx = p.context()
        st = p.stats(); tm = ctx.timings()
        n_cand = int(st.n_candidates)
        d2h = n_cand * 46 + 3 * 4 * (n_cand + 1) + 2 * int(st.n_fragments)  # candidate columns + list offsets + labels (lists add ~4 B per supporting read)
        dev_ms = tm.read_filters_ms + tm.find_fusions_ms + tm.merge_adjacent_ms + tm.evalue_ms + tm.kmer_index_ms + tm.homologs_ms + tm.mismappers_ms
        res = (int(st.n_fragments), e2e_s, dev_ms, st, tm, d2h, n_cand)
        if rank == 0:   # progress on stderr: a run that is cut off still says where the time went
            ev = {n: round(st.event_seconds[i], 2) for i, n in enumerate(L.EV_NAMES) if st.event_seconds[i] >= 0.5}
            print("[bench] step: e2e %.2f s, device %.1f ms, ingest %.2f, annotate %.2f, upload %.2f, output %.2f, events >= 0.5 s: %s" %
                  (e2e_s, dev_ms, st.seconds[L.STEP_INGEST], st.seconds[L.STEP_ANNOTATE], st.seconds[L.STEP_UPLOAD], st.output_seconds, ev), file=sys.stderr, flush=True)
        p.close()
        return res

    sharded = args.mode == "sharded" and world > 1
    if sharded:
        config["sharding"] = "ONE sample, fragments partitioned by contig pair over the ranks (LPT), two NCCL all-gathers (labels, candidates); every rank decodes the BAM"
    launches0 = L.load().arb_kernel_launches()
    for _ in range(args.warmup):
        one_step(sharded)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank); sampler.start()
    launches1 = L.load().arb_kernel_launches()
    t_begin = time.perf_counter()
    results = [one_step(sharded) for _ in range(args.steps)]
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    wall = time.perf_counter() - t_begin
    extra_sharded = None
    slowest = torch.tensor([max(r[1] for r in results)], device="cuda", dtype=torch.float64)
    if dist:
        dist.all_reduce(slowest, op=dist.ReduceOp.MAX)
    if dist and not sharded and not args.no_sharded_extra and float(slowest[0]) < 60.0:   # the same sample once more as ONE job over all ranks: exercises the two all-gathers (skipped when a step takes more than a minute)
        r = one_step(True)
        t = torch.tensor([r[1]], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        extra_sharded = {"e2e_seconds": float(t[0]), "e2e_value": r[0] / float(t[0]), "unit": UNIT, "scaling": "strong",
                         "note": "one sample partitioned by contig pair over all ranks, 2 NCCL all-gathers; host decode is replicated, so this mode buys device memory and device time, not host time"}
    sampler.stop_flag = True; sampler.join(timeout=2)
    launches = L.load().arb_kernel_launches() - launches1
    n_frag = results[0][0]
    dev_ms = sum(r[2] for r in results) / len(results)
    e2e_s = sum(r[1] for r in results) / len(results)
    cls_ms = sum(r[4].cascade_head_ms + r[4].cascade_sequences_ms for r in results) / len(results)
    if dist:
        t = torch.tensor([dev_ms, e2e_s, cls_ms, wall], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, cls_ms, wall = [float(x) for x in t.tolist()]
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    st, tm = results[-1][3], results[-1][4]
    peak, peak_src = measured_peak()
    cascade_bytes = int(tm.cascade_algorithmic_bytes[0]) + int(tm.cascade_algorithmic_bytes[1])   # of the fragments the two launches actually saw (a shard, in sharded mode)
    achieved = cascade_bytes / (cls_ms * 1e-3) / 1e9
    traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum of the two cascade launches on this workload, from the committed `ncu --set full` capture
    try:
        t = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")))
        traffic = t.get(args.workload, {}).get("cascade_dram_bytes_per_step")
    except Exception:
        pass
    jobs = 1 if sharded else world
    line = {"metric": metric, "value": n_frag * jobs / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
            "clocks": sampler.summary(),
            "e2e": {"value": n_frag * jobs / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(tm.h2d_bytes), "d2h_bytes_per_step": int(results[-1][5]),
                    "seconds_per_step": e2e_s, "host_seconds": {n: round(st.seconds[i], 3) for i, n in enumerate(L.STEP_NAMES) if i > 0},
                    "event_seconds": {n: round(st.event_seconds[i], 3) for i, n in enumerate(L.EV_NAMES) if st.event_seconds[i] >= 0.001}, "output_seconds": round(st.output_seconds, 3),
                    "ingest_split": {"inflate": round(st.t_inflate, 3), "parse": round(st.t_parse, 3), "finalize": round(st.t_finalize, 3)}},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "read-level cascade: k_for_each<cascade_head_fn> + k_for_each_scratch<cascade_sequences_fn>", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": cascade_bytes, "kernel_ms": cls_ms,
                         "cascade": {"head_ms": tm.cascade_head_ms, "sequences_ms": tm.cascade_sequences_ms, "queued": int(tm.cascade_queued),
                                     "head_bytes": int(tm.cascade_algorithmic_bytes[0]), "sequences_bytes": int(tm.cascade_algorithmic_bytes[1])},
                         "device_ms": {"duplicates": tm.duplicates_ms, "classify": tm.classify_ms, "read_filters_total": tm.read_filters_ms, "find_fusions_total": tm.find_fusions_ms, "h2d": tm.h2d_ms,
                                       "merge_adjacent": tm.merge_adjacent_ms, "evalue": tm.evalue_ms, "kmer_index": tm.kmer_index_ms, "homologs": tm.homologs_ms, "mismappers": tm.mismappers_ms,
                                       "mismappers_pass1": tm.mismappers_pass1_ms, "mismappers_pass2": tm.mismappers_pass2_ms},
                         "mismapper_items": int(tm.mismapper_items), "mismapper_heavy_items": int(tm.mismapper_heavy_items), "mismapper_tasks": int(tm.mismapper_tasks), "mismapper_rounds": int(tm.mismapper_rounds), "mismapper_registry": {"slots": int(tm.mismapper_table_slots), "overflow": int(tm.mismapper_overflow)}, "kmer_positions": int(tm.kmer_positions)},
            "candidates": int(results[-1][6]), "unfiltered_candidates": int(st.n_unfiltered_candidates), "fragments_per_step": n_frag, "wall_seconds_timed_region": wall}
    if extra_sharded:
        line["sharded_single_sample"] = extra_sharded
    if not args.no_cpu_baseline and world == 1:   # the CPU reference beside the GPU number: rank 0 at N=1 only
        sp = ensure_world(args.workload, sample_bp)
        n, scope_s, total = reference_run(sp, cores)
        line["cpu_baseline"] = {"value": n / scope_s, "unit": UNIT, "cores": 1, "kind": "reference",
                                "sample": "first %d breakpoints of the workload at full depth = %d fragments; unmodified reference (oracle/_ref/arriba), its own time stamps over %s" % (sample_bp, n, SCOPE),
                                "whole_run_seconds": total, "host_cores_available": cores}
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
