"""One sample on several ranks (gloo, CPU stand-in library): rank 0's outputs must be byte-identical to the single-rank run; the work must really be divided."""
import os
import subprocess
import sys
import numpy as np
import pytest
import worldutil
from arriba_b200 import lib as L, _build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json
sys.path.insert(0, {root!r})
import torch.distributed as dist
from arriba_b200 import lib as L, sharded
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
if {backend!r} == "nccl":
    import torch; torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group({backend!r})
out = os.path.join({outdir!r}, "rank%d" % rank); os.makedirs(out, exist_ok=True)
p = L.Pipeline({prefix!r} + ".bam", {prefix!r} + ".gtf", {prefix!r} + ".fa", threads=2, device=int(os.environ.get("LOCAL_RANK", "0")) if {backend!r} == "nccl" else 0, lib_path={lib!r},
               output=os.path.join(out, "fusions.tsv"), discarded=os.path.join(out, "fusions.discarded.tsv"))
sharded.run_sharded(p, rank, world)
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
    return open(os.path.join(outdir, "rank0", "fusions.tsv"), "rb").read(), open(os.path.join(outdir, "rank0", "fusions.discarded.tsv"), "rb").read()


def single(world_obj, lib_path, tmp_path):
    out = str(tmp_path / "single"); os.makedirs(out, exist_ok=True)
    p = L.Pipeline(world_obj.prefix + ".bam", world_obj.prefix + ".gtf", world_obj.prefix + ".fa", threads=2, lib_path=lib_path,
                   output=os.path.join(out, "fusions.tsv"), discarded=os.path.join(out, "fusions.discarded.tsv"))
    p.run_all(); p.close()
    return open(os.path.join(out, "fusions.tsv"), "rb").read(), open(os.path.join(out, "fusions.discarded.tsv"), "rb").read()


def test_partition_is_closed_and_balanced(worlds, hostsim_lib):
    """Fragments that share a duplicate key's contig pair or a candidate's contig pair must share a part; no part may be left (nearly) empty."""
    w = worlds.get("small")
    p = L.Pipeline(w.prefix + ".bam", w.prefix + ".gtf", w.prefix + ".fa", threads=2, lib_path=hostsim_lib)
    for s in (L.STEP_LOAD_REFERENCE, L.STEP_INGEST, L.STEP_ANNOTATE):
        p.step(s)
    fr = p.fragments()
    n = fr["n_fragments"]
    parts = 4
    keys, owner_of_key = p.work_partition(parts)
    p.close()
    assert np.all(np.diff(keys.astype(np.int64)) > 0)
    c = fr["contig"].reshape(3, n).astype(np.int64); split = fr["n_aln"] == 3
    pair = lambda a, b: np.minimum(a, b) << 16 | np.maximum(a, b)
    cand = np.where(split, pair(c[1], c[2]), pair(c[0], c[1]))
    dup = np.where(split, pair(c[0], c[2]), pair(c[0], c[1]))
    pos = np.searchsorted(keys, cand); assert np.all(keys[pos] == cand)
    owner = owner_of_key[pos].astype(np.int64)
    posd = np.searchsorted(keys, dup); assert np.all(keys[posd] == dup)
    assert np.all(owner_of_key[posd] == owner), "a duplicate key's contig pair lives on another part than the fragment"
    assert len(np.unique(cand)) >= parts
    sizes = np.bincount(owner, minlength=parts)
    assert sizes.min() > 0.5 * n / parts, sizes


@pytest.mark.parametrize("n_ranks", [2, 3, 8])   # 8 = the largest world the scaling run uses
def test_sharded_equals_single_hostsim(worlds, hostsim_lib, tmp_path, n_ranks):
    w = worlds.get("cfg5", **dict(scale=0.002, genes=800, breakpoints=400, fragments=40000, extra=("--mismapper-frac", "0.3", "--paralog-frac", "0.15")))
    want = single(w, hostsim_lib, tmp_path)
    got = run_world(w, hostsim_lib, n_ranks, "gloo", tmp_path)
    assert got[0] == want[0], "fusions.tsv differs from the single-rank run"
    assert got[1] == want[1], "fusions.discarded.tsv differs from the single-rank run"


@pytest.mark.gpu
def test_sharded_two_ranks_one_gpu_cuda(worlds, cuda_lib, tmp_path):
    """Two ranks on ONE GPU (gloo carries the exchanges, staged through the host: NCCL refuses two ranks per device): the whole multi-GPU protocol --
    replication of the resident state, divided find_fusions, candidate all-gather + device merge, divided re-alignment -- on the CUDA library."""
    w = worlds.get("cfg5", **dict(scale=0.002, genes=800, breakpoints=400, fragments=40000, extra=("--mismapper-frac", "0.3", "--paralog-frac", "0.15")))
    want = single(w, cuda_lib, tmp_path)
    got = run_world(w, cuda_lib, 2, "gloo", tmp_path)
    assert got == want


@pytest.mark.gpu
def test_sharded_nccl_cuda(worlds, cuda_lib, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (the one-GPU box runs the two-ranks-on-one-GPU test instead)")
    w = worlds.get("cfg5", **dict(scale=0.002, genes=800, breakpoints=400, fragments=40000, extra=("--mismapper-frac", "0.3", "--paralog-frac", "0.15")))
    want = single(w, cuda_lib, tmp_path)
    got = run_world(w, cuda_lib, min(4, torch.cuda.device_count()), "nccl", tmp_path)
    assert got == want
