"""One sample on several GPUs: the two exchanges of the sharded run (include/arriba_b200.h, arriba_b200/csrc/host/shard.cpp) over torch.distributed.

Each rank is one process bound to one GPU (`torchrun`); NCCL carries the blobs over NVLink/NVSwitch (gloo in the CPU test-suite). The library itself is
transport-agnostic: it exports a byte blob per exchange and imports the blobs of all ranks."""
import numpy as np

from . import lib as L


def all_gather_bytes(blob):
    """All-gather of one ragged byte buffer per rank: sizes first, then the payload padded to the longest."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    size = torch.tensor([blob.size], dtype=torch.int64, device=dev)
    sizes = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(sizes, size)
    sizes = [int(s) for s in sizes.tolist()]
    longest = max(sizes)
    mine = torch.zeros(longest, dtype=torch.uint8, device=dev)
    mine[:blob.size] = torch.from_numpy(blob).to(dev)
    everything = torch.empty(world * longest, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(everything, mine)
    host = everything.cpu().numpy()
    return [host[r * longest:r * longest + sizes[r]] for r in range(world)]


def run_sharded(pipeline, rank, world, last_event=None, write_output=True, gather=all_gather_bytes, reference_loaded=False):
    """Runs `pipeline` (an arriba_b200.lib.Pipeline) as rank `rank` of `world`; every rank ends with the complete result, rank 0 writes the files."""
    pipeline.plan_shard(world)
    for s in (L.STEP_LOAD_REFERENCE, L.STEP_INGEST, L.STEP_ANNOTATE):
        if s == L.STEP_LOAD_REFERENCE and reference_loaded:
            continue
        pipeline.step(s)
    pipeline.set_shard(rank, world)
    pipeline.step(L.STEP_UPLOAD)
    pipeline.step(L.STEP_READ_FILTERS)
    if world > 1:
        pipeline.import_shards(L.EXCHANGE_LABELS, gather(pipeline.export_shard(L.EXCHANGE_LABELS)))
    pipeline.step(L.STEP_FRAGMENT_LENGTH)
    pipeline.step(L.STEP_FIND_FUSIONS)
    if world > 1:
        pipeline.import_shards(L.EXCHANGE_CANDIDATES, gather(pipeline.export_shard(L.EXCHANGE_CANDIDATES)))
    pipeline.events(len(L.EV_NAMES) - 1 if last_event is None else last_event)
    if write_output and rank == 0:
        pipeline.write_output()
