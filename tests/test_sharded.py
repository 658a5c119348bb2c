"""One sample on two ranks (gloo, CPU stand-in library): outputs must be byte-identical to the single-rank run and to the oracle."""
import os
import subprocess
import sys
import numpy as np
import pytest
import worldutil
from arriba_b200 import lib as L, _build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch.distributed as dist
from arriba_b200 import lib as L, sharded
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group({backend!r})
if {backend!r} == "nccl":
    import torch; torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
out = os.path.join({outdir!r}, "rank%d" % rank); os.makedirs(out, exist_ok=True)
p = L.Pipeline({prefix!r} + ".bam", {prefix!r} + ".gtf", {prefix!r} + ".fa", threads=2, device=int(os.environ.get("LOCAL_RANK", "0")) if {backend!r} == "nccl" else 0, lib_path={lib!r},
               output=os.path.join(out, "fusions.tsv"), discarded=os.path.join(out, "fusions.discarded.tsv"))
sharded.run_sharded(p, rank, world, write_output=False)
p.write_output()   # every rank writes: all of them must hold the complete, identical result
p.close()
dist.destroy_process_group()
"""


def run_world(world_obj, lib_path, n_ranks, backend, tmp_path):
    outdir = str(tmp_path / ("w%d" % n_ranks)); os.makedirs(outdir, exist_ok=True)
    script = os.path.join(outdir, "worker.py")
    open(script, "w").write(WORKER.format(root=ROOT, backend=backend, outdir=outdir, prefix=world_obj.prefix, lib=lib_path))
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, "tests"))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n_ranks), "--master-addr", "127.0.0.1", "--master-port", str(29500 + n_ranks + os.getpid() % 200), script]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    return [(open(os.path.join(outdir, "rank%d" % k, "fusions.tsv"), "rb").read(), open(os.path.join(outdir, "rank%d" % k, "fusions.discarded.tsv"), "rb").read()) for k in range(n_ranks)]


def single(world_obj, lib_path, tmp_path):
    out = str(tmp_path / "single"); os.makedirs(out, exist_ok=True)
    p = L.Pipeline(world_obj.prefix + ".bam", world_obj.prefix + ".gtf", world_obj.prefix + ".fa", threads=2, lib_path=lib_path,
                   output=os.path.join(out, "fusions.tsv"), discarded=os.path.join(out, "fusions.discarded.tsv"))
    p.run_all(); p.close()
    return open(os.path.join(out, "fusions.tsv"), "rb").read(), open(os.path.join(out, "fusions.discarded.tsv"), "rb").read()


def test_partition_is_closed_and_balanced(worlds, hostsim_lib):
    """Fragments that share a duplicate key's contig pair or a candidate's contig pair must share a rank; no rank may be left (nearly) empty."""
    w = worlds.get("small")
    p = L.Pipeline(w.prefix + ".bam", w.prefix + ".gtf", w.prefix + ".fa", threads=2, lib_path=hostsim_lib)
    for s in (L.STEP_LOAD_REFERENCE, L.STEP_INGEST, L.STEP_ANNOTATE):
        p.step(s)
    fr = p.fragments()
    n = fr["n_fragments"]
    world = 4
    p.set_shard(1, world)
    owner = np.full(n, -1)
    for r in range(world):
        m = p.shard_members(r)
        assert np.all(np.diff(m.astype(np.int64)) > 0)
        assert np.all(owner[m] == -1)
        owner[m] = r
    p.close()
    assert np.all(owner >= 0)
    c = fr["contig"].reshape(3, n).astype(np.int64); split = fr["n_aln"] == 3
    pair = lambda a, b: np.minimum(a, b) << 16 | np.maximum(a, b)
    cand = np.where(split, pair(c[1], c[2]), pair(c[0], c[1]))
    dup = np.where(split, pair(c[0], c[2]), pair(c[0], c[1]))
    for key in (cand, dup):
        order = np.argsort(key, kind="stable")
        same = key[order][1:] == key[order][:-1]
        assert np.all(owner[order][1:][same] == owner[order][:-1][same])
    assert len(np.unique(cand)) >= world
    sizes = np.bincount(owner, minlength=world)
    assert sizes.min() > 0.5 * n / world, sizes


@pytest.mark.parametrize("n_ranks", [2, 3])
def test_sharded_equals_single_hostsim(worlds, hostsim_lib, tmp_path, n_ranks):
    w = worlds.get("cfg5", **dict(scale=0.002, genes=800, breakpoints=400, fragments=40000, extra=("--mismapper-frac", "0.3", "--paralog-frac", "0.15")))
    want = single(w, hostsim_lib, tmp_path)
    got = run_world(w, hostsim_lib, n_ranks, "gloo", tmp_path)
    for k in range(n_ranks):
        assert got[k][0] == want[0], "fusions.tsv of rank %d differs from the single-rank run" % k
        assert got[k][1] == want[1], "fusions.discarded.tsv of rank %d differs from the single-rank run" % k


@pytest.mark.gpu
def test_sharded_equals_single_cuda(worlds, cuda_lib, tmp_path):
    """World size 1 on the GPU box exercises export -> import on the CUDA library (the multi-GPU run itself is bench.py --gpus N --mode sharded)."""
    import torch
    w = worlds.get("small")
    want = single(w, cuda_lib, tmp_path)
    n = min(2, torch.cuda.device_count())
    got = run_world(w, cuda_lib, n, "nccl" if n > 1 else "gloo", tmp_path)
    for k in range(n):
        assert got[k] == want
