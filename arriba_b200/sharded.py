"""One sample on several GPUs: the launcher-side driver (include/arriba_b200.h, "one sample on several GPUs"; arriba_b200/csrc/exchange.cu).

Each rank is one process bound to one GPU (`torchrun`). The library exposes DEVICE buffers; this module moves them with torch.distributed -- NCCL over
NVLink/NVSwitch on the GPU box, gloo on host memory in the CPU test-suite (the stand-in library's "device" buffers are host memory) -- and never stages
them through the host.

  rank 0     ingests the BAM, annotates on its GPU, then BROADCASTS the genome, the annotation and the resident fragment table (exchange groups);
  every rank runs the read-level cascade on its replica (20 ms, no exchange), then find_fusions for the contig pairs it owns;
  all-gather of the packed candidate tables (ONE collective), every rank merges them in first-insertion order; mate swaps combined by a MAX all-reduce;
  rank 0     runs the event-level chain; at filter_mismappers it broadcasts the stage's inputs, every rank re-aligns every W-th work item, the verdict
             bytes are combined by a MAX all-reduce, rank 0 finishes the stage and the chain and writes the files."""
import ctypes as C

import numpy as np

from . import lib as L

XG_CONTIGS, XG_ANNOTATION, XG_TABLE, XG_MISMAP_STATE = 0, 1, 2, 3


class _CudaBytes:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def view(ptr, nbytes, cuda):
    """torch uint8 tensor over `nbytes` bytes of library-owned memory at `ptr` (device memory of the CUDA library, host memory of the stand-in); no copy."""
    import torch
    if nbytes == 0 or not ptr:
        return torch.empty(0, dtype=torch.uint8, device="cuda" if cuda else "cpu")
    if cuda:
        return torch.as_tensor(_CudaBytes(ptr, nbytes), device="cuda")
    return torch.from_numpy(np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(ptr)))


class Transport:
    """Collectives over library-owned buffers. NCCL works on the device buffers in place. With the gloo backend and the CUDA library (the test that runs two
    ranks on ONE GPU, which NCCL refuses) the buffers are staged through host tensors around each collective; with the stand-in library they are host memory."""
    def __init__(self, rank, world, lib_is_cuda):
        import torch
        import torch.distributed as dist
        self.rank, self.world, self.dist, self.torch = rank, world, dist, torch
        self.cuda = lib_is_cuda                       # the library's buffers are device memory
        self.direct = dist.get_backend() == "nccl" or not lib_is_cuda   # collectives run on the buffers themselves
        self.dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")

    def fence(self):
        """The collectives run on the transport's own CUDA stream, the library's kernels on the context's: the host waits for the device before the library
        touches a buffer a collective wrote (and before a buffer a collective still reads can be recycled)."""
        if self.cuda:
            self.torch.cuda.synchronize()

    def _bcast(self, v):
        if self.direct:
            self.dist.broadcast(v, 0)
        else:
            h = v.cpu(); self.dist.broadcast(h, 0); v.copy_(h)

    def bcast_ints(self, values, n):
        """list of n integers from rank 0"""
        t = self.torch.zeros(n, dtype=self.torch.int64, device=self.dev)
        if self.rank == 0:
            t[:len(values)] = self.torch.tensor([int(v) for v in values], dtype=self.torch.int64)
        self.dist.broadcast(t, 0)
        return [int(x) for x in t.tolist()]

    def bcast_array(self, a, dtype):
        """numpy array from rank 0 (size first)"""
        n = self.bcast_ints([a.size] if self.rank == 0 else [], 1)[0]
        t = self.torch.zeros(max(n, 1), dtype=self.torch.int64, device=self.dev)
        if self.rank == 0 and n:
            t[:n] = self.torch.from_numpy(np.ascontiguousarray(a).astype(np.int64))
        self.dist.broadcast(t, 0)
        return t[:n].cpu().numpy().astype(dtype)

    def bcast_group(self, ctx, group):
        """replicates one exchange group of rank 0's context on every rank, device to device"""
        header = ctx.exchange_header(group) if self.rank == 0 else []
        n = self.bcast_ints([len(header)], 1)[0]
        header = self.bcast_ints(header, n)
        if self.rank != 0:
            ctx.exchange_prepare(group, header)
        for p, b in ctx.exchange_buffers(group):
            if b:
                self._bcast(view(p, b, self.cuda))
        self.fence()
        if self.rank != 0:
            ctx.exchange_commit(group)

    def max_bytes(self, ptr, nbytes):
        """in-place MAX all-reduce over one byte buffer per rank"""
        if nbytes:
            v = view(ptr, nbytes, self.cuda)
            if self.direct:
                self.dist.all_reduce(v, op=self.dist.ReduceOp.MAX)
            else:
                h = v.cpu(); self.dist.all_reduce(h, op=self.dist.ReduceOp.MAX); v.copy_(h)
            self.fence()

    def gather_candidates(self, ctx):
        """the single all-gather of the candidate tables, then the merge on every rank"""
        torch, dist = self.torch, self.dist
        blob, nbytes, sizes = ctx.candidates_export()
        mine = torch.tensor(sizes + [nbytes], dtype=torch.int64, device=self.dev)
        everyone = torch.empty(5 * self.world, dtype=torch.int64, device=self.dev)
        dist.all_gather_into_tensor(everyone, mine)
        everyone = [int(x) for x in everyone.tolist()]
        stride = (max(everyone[5 * r + 4] for r in range(self.world)) + 255) // 256 * 256
        send = torch.zeros(stride, dtype=torch.uint8, device=self.dev)
        send[:nbytes] = view(blob, nbytes, self.cuda)
        recv = torch.empty(stride * self.world, dtype=torch.uint8, device=self.dev)
        dist.all_gather_into_tensor(recv, send)
        if self.cuda and not self.direct:
            recv = recv.cuda()
        self.fence()
        ctx.candidates_import(recv.data_ptr(), stride, [everyone[5 * r + k] for r in range(self.world) for k in range(4)], self.world)
        self.fence()   # `recv` is released when this returns


def run_sharded(pipeline, rank, world, last_event=None, write_output=True, reference_loaded=False):
    """Runs `pipeline` (an arriba_b200.lib.Pipeline) as rank `rank` of `world`; rank 0 holds the result and writes the files."""
    if world == 1:
        for s in range(L.STEP_LOAD_REFERENCE if not reference_loaded else L.STEP_INGEST, L.STEP_COUNT):
            pipeline.step(s)
        pipeline.events(len(L.EV_NAMES) - 1 if last_event is None else last_event)
        if write_output:
            pipeline.write_output()
        return
    tr = Transport(rank, world, pipeline.lib.arb_backend().decode().startswith("cuda"))
    import os, sys, time
    trace = os.environ.get("ARB_TRACE") and rank == 0
    t_last = [time.perf_counter()]
    def lap(what):
        if trace:
            tr.fence(); t = time.perf_counter(); sys.stderr.write("[laps] sharded                %-34s %8.1f ms\n" % (what, (t - t_last[0]) * 1e3)); t_last[0] = t
    last_event = len(L.EV_NAMES) - 1 if last_event is None else last_event
    if rank == 0:
        for s in (L.STEP_LOAD_REFERENCE, L.STEP_INGEST, L.STEP_ANNOTATE, L.STEP_UPLOAD):
            if s == L.STEP_LOAD_REFERENCE and reference_loaded:
                continue
            pipeline.step(s)
    else:
        pipeline.attach_device()
    ctx = pipeline.context()
    lap("ingest + annotate (rank 0)")
    for g in (XG_CONTIGS, XG_ANNOTATION, XG_TABLE):
        tr.bcast_group(ctx, g)
    lap("genome, annotation, table -> all ranks")
    keys, owner = pipeline.work_partition(world) if rank == 0 else (np.zeros(0, np.uint32), np.zeros(0, np.uint8))
    keys = tr.bcast_array(keys, np.uint32); owner = tr.bcast_array(owner, np.uint8)
    ctx.set_work_partition(keys, owner, rank, world)
    lap("work partition")
    if rank == 0:
        pipeline.step(L.STEP_READ_FILTERS); pipeline.step(L.STEP_FRAGMENT_LENGTH)
        gap = int(pipeline.stats().max_mate_gap)
    else:
        ctx.run_read_filters(); gap = 0
    gap = tr.bcast_ints([gap], 1)[0]
    lap("read filters + fragment length")
    if rank == 0:
        pipeline.step(L.STEP_FIND_FUSIONS)
    else:
        ctx.find_fusions(gap)
    lap("find_fusions (own contig pairs)")
    tr.gather_candidates(ctx)
    p, n = ctx.swaps_buffer(); tr.max_bytes(p, n); ctx.swaps_apply()
    lap("candidate all-gather + merge + mate swaps")
    if last_event < L.EV_NAMES.index("mismappers"):
        if rank == 0:
            pipeline.events(last_event)
        return
    active = pipeline.mismappers_begin() if rank == 0 else False
    active = bool(tr.bcast_ints([int(active)], 1)[0])
    lap("event chain up to the re-alignment (rank 0)")
    if active:
        tr.bcast_group(ctx, XG_MISMAP_STATE)
        lap("re-alignment inputs -> all ranks")
        p, n = ctx.filter_mismappers_part(gap, rank, world)
        tr.max_bytes(p, n)
        if rank == 0:
            ctx.filter_mismappers_finish()
        lap("re-alignment (own items) + verdict all-reduce")
    if rank == 0:
        pipeline.mismappers_end()
        pipeline.events(last_event)
        lap("rest of the event chain (rank 0)")
        if write_output:
            pipeline.write_output()
        lap("output (rank 0)")
