"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch).

The path shards like this: points REPLICATED on every GPU, hypotheses PARTITIONED
by global index (rank r owns [r*H/W, (r+1)*H/W); Philox counters and lexicographic ranks are
global, so results do not depend on W), and exactly two exchange steps:
  1. one all-reduce(MAX) on the packed 64-bit key (count << 32) | (0xFFFFFFFF - index): largest
     count wins, ties go to the smallest index -- the strict '>' of RANSAC.hxx:100,245;
  2. one all-reduce(SUM) of <= 91 doubles per refine pass (per LM evaluation for the iterative fits):
     the least-squares moments of the point shards.
Both run in place on device memory through the hooks of lsqr_set_shard.  Replicating the points is the one bulk
transfer: `upload_replicated` moves 1/W of the host buffer over each GPU's PCIe link and lets an NCCL all-gather over
NVLink / NVSwitch fan it out, instead of W full uploads competing for the host's memory bandwidth.  The helpers below are
pure functions so that the N>1 logic is testable with gloo on CPU.
"""
import contextlib
import ctypes

import numpy as np

INDEX_MASK = 0xFFFFFFFF


def hypothesis_shard(count, rank, world):
    """[lo, hi) of the hypotheses of one request owned by `rank` (same formula as engine.cu)."""
    return count * rank // world, count * (rank + 1) // world


def point_shard(n, rank, world):
    """[begin, end) of the data refined by `rank`; boundaries fall on 32-datum words (engine.cu shard_range)."""
    if world <= 1:
        return 0, n
    words = (n + 31) // 32
    return min(words * rank // world * 32, n), min(words * (rank + 1) // world * 32, n)


def pack_key(count, index):
    return (int(count) << 32) | (INDEX_MASK - int(index))


def unpack_key(key):
    return int(key) >> 32, INDEX_MASK - (int(key) & INDEX_MASK)


def best_key(counts, index_base=0):
    """arg-max of a count vector as a packed key (first maximum wins)."""
    counts = np.asarray(counts)
    if counts.size == 0:
        return 0
    i = int(np.argmax(counts))  # numpy returns the first maximum
    return pack_key(counts[i], index_base + i)


class _DevArray:
    """Wraps a raw device pointer for torch.as_tensor via __cuda_array_interface__."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def install_hooks(engine, rank, world, group=None):
    """Registers NCCL all-reduce hooks on an Engine.  The keys are < 2^63 (counts < 2^31), so the
    MAX over int64 equals the MAX over uint64."""
    import torch
    import torch.distributed as dist

    def on(stream):
        # the hook contract is "stream-ordered on the stream handed in": make it torch's current stream for the collective, so
        # that NCCL's own stream waits for it and it waits for NCCL (a null handle means the legacy default stream)
        return torch.cuda.stream(torch.cuda.ExternalStream(int(stream))) if stream else contextlib.nullcontext()

    def max_hook(user, dev_key, stream):
        try:
            with on(stream):
                t = torch.as_tensor(_DevArray(dev_key, 1, "<i8"), device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            return 0
        except Exception as e:  # never raise across the C ABI
            print("max all-reduce hook failed:", e)
            return 1

    def sum_hook(user, dev_vals, count, stream):
        try:
            with on(stream):
                t = torch.as_tensor(_DevArray(dev_vals, count, "<f8"), device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            return 0
        except Exception as e:
            print("sum all-reduce hook failed:", e)
            return 1

    engine.set_shard(rank, world, max_hook, sum_hook)


def upload_slice(n, rank, world):
    """[begin, end) of the records that `rank` copies from the host, and the padded slice length used by the all-gather."""
    per = (n + world - 1) // world
    return min(rank * per, n), min((rank + 1) * per, n), per


_GATHER_CACHE = {}


def gather_replicated(host, rank, world, device, group=None):
    """host: [n, D] float64 tensor holding (at least) this rank's slice.  Returns the full [n, D] tensor on `device`:
    own slice copied host -> device, the rest by all-gather."""
    import torch
    import torch.distributed as dist

    n, d = host.shape
    lo, hi, per = upload_slice(n, rank, world)
    key = (str(device), per * world, d)
    full = _GATHER_CACHE.get(key)
    if full is None:
        full = _GATHER_CACHE[key] = torch.empty((per * world, d), dtype=torch.float64, device=device)
    mine = full[rank * per:(rank + 1) * per]
    if hi > lo:
        mine[: hi - lo].copy_(host[lo:hi], non_blocking=True)
    if world > 1:
        dist.all_gather_into_tensor(full, mine, group=group) if full.is_cuda else \
            dist.all_gather(list(full.view(world, per, d).unbind(0)), mine.clone(), group=group)
    return full[:n]


def upload_replicated(engine, host, rank, world, group=None):
    """RANSAC::compute's `data` argument for W ranks that all need every point: 1/W of the bytes per PCIe link, the rest
    over NVLink.  `host` is a pinned [n, D] float64 torch tensor (packed records)."""
    import torch

    full = gather_replicated(host, rank, world, torch.device("cuda", torch.cuda.current_device()), group)
    engine.upload_device(full.data_ptr(), host.shape[0])
    return full


def full_mask(mask, group=None):
    """In sharded mode a rank writes only its own point shard of the consensus set (lsqr_b200.h, lsqr_set_shard): this
    assembles the set RANSAC::compute returns on every rank.  `mask`: uint8[n] from Engine.get_mask() / Engine.ransac()."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return mask
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = point_shard(len(mask), rank, world)
    mine = np.zeros_like(mask)
    mine[lo:hi] = mask[lo:hi]
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.from_numpy(mine).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t.cpu().numpy()


def init_native(engine, rank, world, group=None):
    """One process per GPU with the library's own NCCL communicator: rank 0 creates the unique id, torch.distributed only
    carries its 128 bytes to the other ranks; every collective of the path is then an ncclAllReduce / ncclAllGather issued
    by liblsqr_b200.so on its own stream (lsqr_ctx_init_nccl)."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return
    ident = [engine.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0, group=group)
    engine.init_nccl(ident[0], rank, world)
