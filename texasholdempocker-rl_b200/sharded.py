"""Multi-GPU sharding: one process per GPU, work split by sample index / combination index, one
all-reduce(SUM) of the integer counts (NCCL over NVLink on the GPU box, gloo in the CPU tests).

Integer sums make the result bit-identical for any world size and any reduction order.
"""
import numpy as np

from . import _native as N

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def shard_range(total, rank, world):
    """Contiguous shard [begin, end) of range(total) for `rank` of `world`; shards tile the range exactly."""
    total, rank, world = int(total), int(rank), int(world)
    return total * rank // world, total * (rank + 1) // world


def _world(group):
    if dist is None or not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def sharded_equity_mc(states, n_opp, rollouts, seed, sample_offset=0, group=None, local_fn=None, reduce=True):
    """Every rank holds the same `states`; rank r rolls samples shard_range(rollouts, r, world) and the
    (win, tie, loss) counts are all-reduced.  local_fn(states, n_opp, rollouts, seed, sample_offset=...)
    defaults to the CUDA path (tests inject the oracle there).

    Host arrays in -> host array out on every rank: with the NCCL backend the states go to the device once
    (H2D), the shard is rolled there, the int64[S,3] counts are all-reduced over NVLink and copied back (D2H)."""
    from .equity import equity_mc
    fn = local_fn or equity_mc
    rank, world = _world(group)
    b, e = shard_range(rollouts, rank, world)
    on_host = not (torch is not None and isinstance(states, torch.Tensor))
    if on_host and local_fn is None and world > 1 and reduce and dist.get_backend(group) == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
        t = torch.from_numpy(np.ascontiguousarray(states, dtype=np.int64)).to(dev, non_blocking=True)
        counts = equity_mc(t, n_opp, e - b, seed, sample_offset=sample_offset + b)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
        return counts.cpu().numpy()
    counts = fn(states, n_opp, e - b, seed, sample_offset=sample_offset + b)
    if world > 1 and reduce:
        t = counts if (torch is not None and isinstance(counts, torch.Tensor)) else torch.from_numpy(np.ascontiguousarray(counts))
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        counts = t if isinstance(counts, torch.Tensor) else t.numpy()
    return counts


def sharded_enum7(group=None, local_fn=None, comb_begin=0, comb_end=N.NUM_HANDS_7, reduce=True):
    """Rank r enumerates part r of `world` interleaved parts of the colex range (phe_enum7_part: equal cost per
    rank, unlike contiguous slices, because the low indices hold the short, expensive blocks); histograms are
    all-reduced.  local_fn(begin, end) (tests: the oracle) gets a contiguous slice instead."""
    from .evaluator import enumerate_c52_7
    rank, world = _world(group)
    if local_fn is None:
        h9, h = enumerate_c52_7(comb_begin, comb_end, device_out=True, part=rank, nparts=world)
    else:
        b, e = shard_range(comb_end - comb_begin, rank, world)
        h9, h = local_fn(comb_begin + b, comb_begin + e)
    if world > 1 and reduce:
        if not isinstance(h9, torch.Tensor):
            h9, h = torch.from_numpy(np.ascontiguousarray(h9)), torch.from_numpy(np.ascontiguousarray(h))
        dist.all_reduce(h9, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(h, op=dist.ReduceOp.SUM, group=group)
    return h9, h


class FusedEnum7:
    """Exhaustive enumeration sharded over the ranks of `group` with the histogram all-reduce fused into the
    enumeration kernel (phe_enum7_allreduce[_nvls]): the last CTA of each rank's kernel exchanges the 30 KB u32 histogram
    through peer memory over NVLink, so a step is ONE kernel per GPU and can be captured in a CUDA graph.  torch's symmetric
    memory provides the peer-mapped workspace (plumbing only); with one rank no peer mapping is needed.

    nvls: None = use the NVLink-SHARP multicast mapping when symmetric memory offers one and the group has four ranks or more
    (multimem.red.add: one reduction reaches every rank, one multimem.st signals them, no per-rank slots to sum; with two
    ranks plain peer stores measure the same), True = require it, False = plain peer stores.  All ranks decide alike."""

    def __init__(self, group=None, nvls=None):
        import ctypes as C
        self.rank, self.world = _world(group)
        L = N.lib()
        nbytes = int(L.phe_enum7_allreduce_workspace())
        dev = torch.device("cuda", torch.cuda.current_device())
        if self.world > 1:
            import torch.distributed._symmetric_memory as symm
            g = group if group is not None else dist.group.WORLD
            self.ws = symm.empty(nbytes, dtype=torch.uint8, device=dev)
            self.handle = symm.rendezvous(self.ws, g.group_name if hasattr(g, "group_name") else g)
            ptrs = [int(p) for p in self.handle.buffer_ptrs]
            mc = int(getattr(self.handle, "multicast_ptr", 0) or 0)
            have = torch.tensor([1 if mc else 0], device=dev)
            dist.all_reduce(have, op=dist.ReduceOp.MIN, group=group)          # every rank or none
            if nvls is True and not int(have.item()):
                raise N.PheError("FusedEnum7(nvls=True): symmetric memory has no multicast mapping on this system")
            self._mc = mc if (int(have.item()) and (nvls is True or (nvls is None and self.world >= 4))) else 0
        else:
            self.ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            ptrs = [self.ws.data_ptr()]
            self._mc = 0
        self.nvls = bool(self._mc)
        self.ws.zero_()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(group)
        self._ptrs = (C.c_void_p * self.world)(*ptrs)
        self.out = torch.zeros(9 + N.NUM_RANKS + 1, dtype=torch.int64, device=dev)

    def run(self, comb_begin=0, comb_end=N.NUM_HANDS_7):
        """Launches on the current stream; returns views (hist9, hist7463) of the all-reduced result buffer."""
        import ctypes as C
        N.check(N.lib().phe_enum7_allreduce_nvls(C.c_void_p(self.out.data_ptr()), comb_begin, comb_end, self.rank, self.world, self._ptrs,
                                                 C.c_void_p(self._mc) if self._mc else None, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return self.out[:9], self.out[9:]

    def check(self):
        """Synchronises the current stream and raises if a peer did not show up in time for any exchange so far (the output of
        that launch is a partial sum); returns the number of launches completed on this rank."""
        import ctypes as C
        launches, failed = C.c_uint32(0), C.c_uint32(0)
        N.check(N.lib().phe_enum7_allreduce_status(C.c_void_p(self.ws.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream),
                                                   C.byref(launches), C.byref(failed)))
        if failed.value:
            raise N.PheError(f"fused enumeration all-reduce: a peer did not arrive within the time-out at launch {failed.value} "
                             f"(rank {self.rank} of {self.world}); the histogram of that launch is partial")
        return int(launches.value)
