"""Batch sharding across the GPUs of one box: every op on the path is per-sample, so the only
exchange is ONE all-gather of the final depth maps.  One process per GPU (torchrun), NCCL over NVLink;
the same code runs under gloo on CPU for the host-logic tests.  New relative to the reference, whose test
path is single-process nn.DataParallel at batch 1 (reference src/main.py:434)."""
from typing import Callable, Dict, Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, near-equal shards: (first_index, count) of `rank`; earlier ranks take the remainder."""
    if not (0 <= rank < world) or global_batch < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(global_batch, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def slice_sample(sample: Dict[str, torch.Tensor], first: int, count: int) -> Dict[str, torch.Tensor]:
    """Per-rank view of a full-batch sample dict (keys of reference src/data/kittidc.py:273 + optional 'noise')."""
    return {k: (v[first:first + count] if torch.is_tensor(v) and v.dim() > 0 else v) for k, v in sample.items()}


def gather_depth(pred_local: torch.Tensor, global_batch: int, group=None) -> torch.Tensor:
    """The single collective on the path: all ranks end with pred [global_batch,1,H,W].
    Equal shards -> one `all_gather_into_tensor`; ragged shards pad to the largest shard first."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = [shard_range(global_batch, r, world)[1] for r in range(world)]
    cmax = max(counts)
    tail = tuple(pred_local.shape[1:])
    if pred_local.shape[0] != counts[rank]:
        raise ValueError(f"rank {rank} holds {pred_local.shape[0]} maps, expected {counts[rank]}")
    send = pred_local.contiguous()
    if counts[rank] < cmax:
        send = torch.cat([send, send.new_zeros((cmax - counts[rank],) + tail)])
    out = send.new_empty((world * cmax,) + tail)
    dist.all_gather_into_tensor(out, send, group=group)
    if all(c == cmax for c in counts):
        return out
    return torch.cat([out[r * cmax:r * cmax + c] for r, c in enumerate(counts)])


class DepthGatherer:
    """The same single collective, taken OFF the compute stream for steady-state serving: `submit()` enqueues the
    all-gather of this step's maps on a side stream into one of `slots` rotating result buffers and returns at once, so
    the next step's kernels never queue behind a peer that is still finishing the current one (under a power cap the
    ranks' step times differ by a few per cent; a blocking per-step collective makes every rank run at the pace of that
    step's slowest peer — round-1 SCALE: 0.95 at N = 8).  `result(ticket)` makes the current stream wait for that
    gather.  On CPU tensors (gloo tests) it degrades to the synchronous call."""

    def __init__(self, global_batch: int, group=None, slots: int = 2):
        self.global_batch, self.group, self.slots = global_batch, group, max(2, int(slots))
        self._bufs, self._done, self._stream, self._n = [None] * self.slots, [None] * self.slots, None, 0

    def submit(self, pred_local: torch.Tensor) -> int:
        slot = self._n % self.slots
        self._n += 1
        if not pred_local.is_cuda:
            self._bufs[slot] = gather_depth(pred_local, self.global_batch, self.group)
            return slot
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=pred_local.device)
        cur = torch.cuda.current_stream(pred_local.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        self._stream.wait_event(ready)
        with torch.cuda.stream(self._stream):
            self._bufs[slot] = gather_depth(pred_local, self.global_batch, self.group)
            done = torch.cuda.Event()
            done.record(self._stream)
        pred_local.record_stream(self._stream)  # the allocator must not hand these maps out before the gather has read them
        self._done[slot] = done
        return slot

    def result(self, ticket: int) -> torch.Tensor:
        out = self._bufs[ticket]
        if self._done[ticket] is not None:
            cur = torch.cuda.current_stream(out.device)
            cur.wait_event(self._done[ticket])
            out.record_stream(cur)  # allocated on the side stream, consumed here: keep the block until this stream is done
        return out

    def drain(self):
        if self._stream is not None:
            torch.cuda.current_stream(self._stream.device).wait_stream(self._stream)


def run_sharded(model: Callable[[Dict[str, torch.Tensor]], Dict[str, torch.Tensor]], sample: Dict[str, torch.Tensor],
                global_batch: int, group=None) -> torch.Tensor:
    """Each rank runs `model` on its slice of the full-batch `sample`, then one all-gather of 'pred'."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    first, count = shard_range(global_batch, rank, world)
    out = model(slice_sample(sample, first, count))
    return gather_depth(out["pred"], global_batch, group)
