"""Collective plumbing of the contrastive step (replaces x_clip/distributed.py:14-56).

The reference gathers a python LIST of padded fp32 tensors after a size exchange with two host
syncs (distributed.py:18-23, x_clip.py:764) and a second gather for the extra latents
(x_clip.py:767-768).  Here every rank contributes equal-sized shards (asserted), ALL latent sets
travel in ONE contiguous bf16 all-gather, and the only other exchange is a [4, b] fp32 gather of
(lse, positive) pairs so each rank can form the global loss and its backward weights.
Pure torch.distributed (NCCL on GPUs, gloo in the CPU tests); no arithmetic lives here.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when no multi-rank process group exists."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def assert_equal_local_batch(b: int, device: torch.device) -> None:
    """Every rank must contribute the same number of rows (the reference pads to the largest
    batch instead, distributed.py:18-33; here unequal batches are an error, not a hang): one
    tiny MAX all-reduce of (b, -b); the comparison stays on the device (`_assert_async`), so no
    host synchronisation is added to the step."""
    if world()[1] == 1:
        return
    t = torch.tensor([b, -b], device=device, dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = (t[0] + t[1]) == 0                       # max(b) == min(b)
    if device.type == "cuda":
        torch._assert_async(ok, "x_clip_b200: ranks hold different local batch sizes (unsupported; "
                                "use drop_last or pad the last batch)")
    elif not bool(ok):
        raise RuntimeError("x_clip_b200: ranks hold different local batch sizes (unsupported; "
                           "use drop_last or pad the last batch)")


def gather_rows(shards: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """Each shard is this rank's [b, D] block of one latent set.  Returns, per set, the
    [W*b, D] matrix whose rows r*b..(r+1)*b-1 are rank r's - one collective for all sets."""
    rank, w = world()
    if w == 1:
        return [s.contiguous() for s in shards]
    b, D = shards[0].shape
    for s in shards:
        assert tuple(s.shape) == (b, D), "every latent set must have the same local shape"
    assert_equal_local_batch(b, shards[0].device)
    stacked = torch.stack(list(shards)).contiguous()                       # [k, b, D]
    out = torch.empty(w * stacked.numel(), device=stacked.device, dtype=stacked.dtype)
    dist.all_gather_into_tensor(out, stacked.view(-1))                     # flat: backend agnostic
    out = out.view(w, len(shards), b, D)
    return [out[:, j].reshape(w * b, D).contiguous() for j in range(len(shards))]


def gather_stats(stats: torch.Tensor) -> torch.Tensor:
    """stats [k, b] per rank -> [k, W*b] in rank order."""
    rank, w = world()
    if w == 1:
        return stats
    k, b = stats.shape
    out = torch.empty(w * k * b, device=stats.device, dtype=stats.dtype)
    dist.all_gather_into_tensor(out, stats.contiguous().view(-1))
    return out.view(w, k, b).permute(1, 0, 2).reshape(k, w * b).contiguous()


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    if world()[1] > 1:
        dist.all_reduce(t)
    return t


def reduce_scatter_rows(t: torch.Tensor) -> torch.Tensor:
    """t [W*r, D] holds this rank's partial sums for every rank's rows; returns the summed
    [r, D] block of this rank (FILIP: gradients of the gathered image-token latents)."""
    rank, w = world()
    if w == 1:
        return t
    rows = t.shape[0] // w
    t = t.contiguous()
    if dist.get_backend() == "gloo":          # gloo has no reduce_scatter: all-reduce + slice
        dist.all_reduce(t)
        return t[rank * rows:(rank + 1) * rows].clone()
    out = torch.empty((rows,) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
    dist.reduce_scatter_tensor(out, t)
    return out


_active_syncs: List["GradSync"] = []


class defer_grad_sync:
    """`with defer_grad_sync(flag):` - when flag is true, backward passes inside accumulate
    parameter gradients without triggering any live GradSync (used by the micro-batched step,
    engine.ChunkedClipLossFn, whose backward runs one autograd pass per chunk)."""

    def __init__(self, flag: bool = True):
        self.flag = flag
        self.saved = []

    def __enter__(self):
        if self.flag:
            self.saved = [(s, s.enabled) for s in _active_syncs]
            for s in _active_syncs:
                s.enabled = False
        return self

    def __exit__(self, *exc):
        for s, e in self.saved:
            s.enabled = e
        self.saved = []
        return False


class GradSync:
    """Bucketed all-reduce of parameter gradients, launched from inside backward.

    The reference leaves data-parallel weight synchronisation to the user's DDP wrapper
    (its README trains single-GPU).  This is the minimal equivalent
    for the CLIP module of this package: parameters are grouped into buckets in REVERSE
    registration order (late layers finish their backward first); a post-accumulate-grad hook
    counts arrivals and, once a bucket is complete, flattens it and starts an asynchronous
    all-reduce - on NCCL that runs on the communicator's own stream, so it overlaps the rest of
    backward.  `finish()` (call it after `loss.backward()`, before the optimizer) waits for the
    collectives, divides by the world size (`average=True`, the DDP convention) and scatters
    the result back into `p.grad`.

    Every rank must build the same autograd graph (same parameters receive gradients), which
    holds for `CLIP.forward(..., return_loss=True)` with equal flags on all ranks.  Parameters
    that received no gradient (frozen encoder, unused extra projections) are skipped by all
    ranks alike.  With a single rank everything is a no-op.  For gradient accumulation wrap the
    non-final backward passes in `with sync.no_sync():` (as with DDP) - a bucket is reduced once,
    when its last gradient of the pass arrives.
    """

    def __init__(self, module: torch.nn.Module, bucket_bytes: int = 64 << 20, average: bool = True,
                 comm_dtype: torch.dtype | None = None, process_group=None):
        self.world = world()[1]
        self.average = average
        self.comm_dtype = comm_dtype
        self.group = process_group
        self.enabled = True
        self._buckets: List[dict] = []
        self._bucket_of = {}
        self._handles = []
        if self.world == 1:
            return
        params = [p for p in module.parameters() if p.requires_grad]
        cur, cur_bytes = [], 0
        for p in reversed(params):
            nbytes = p.numel() * p.element_size()
            if cur and (cur_bytes + nbytes > bucket_bytes or p.dtype != cur[0].dtype
                        or p.device != cur[0].device):
                self._add_bucket(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self._add_bucket(cur)
        for p in params:
            self._handles.append(p.register_post_accumulate_grad_hook(self._on_grad))
        _active_syncs.append(self)

    def _add_bucket(self, params):
        b = {"params": list(params), "arrived": 0, "work": None, "flat": None, "members": None}
        for p in params:
            self._bucket_of[p] = b
        self._buckets.append(b)

    def no_sync(self):
        """Context manager: backward passes inside it accumulate locally (no collectives)."""
        import contextlib

        @contextlib.contextmanager
        def _cm():
            prev, self.enabled = self.enabled, False
            try:
                yield
            finally:
                self.enabled = prev
        return _cm()

    def _on_grad(self, p):
        if not self.enabled:
            return
        b = self._bucket_of[p]
        if b["work"] is not None or b["members"] is not None:
            raise RuntimeError(
                "GradSync: a parameter received a second gradient after its bucket was reduced - "
                "several backward passes per step must run under `sync.no_sync()` / "
                "`defer_grad_sync()` except the last one")
        b["arrived"] += 1
        if b["arrived"] == len(b["params"]):
            self._launch(b)

    def _launch(self, b):
        members = [p for p in b["params"] if p.grad is not None]
        b["members"] = members
        if not members:
            return
        flat = torch.cat([p.grad.reshape(-1) for p in members])
        if self.comm_dtype is not None and flat.dtype != self.comm_dtype:
            flat = flat.to(self.comm_dtype)
        b["flat"] = flat
        b["work"] = dist.all_reduce(flat, group=self.group, async_op=True)

    def finish(self):
        """Wait for every bucket, write the reduced gradients back, re-arm for the next step."""
        if self.world == 1:
            return
        for b in self._buckets:
            if b["members"] is None:          # some parameter of the bucket got no gradient
                self._launch(b)
        for b in self._buckets:
            if b["work"] is not None:
                b["work"].wait()
                flat = b["flat"]
                off = 0
                for p in b["members"]:
                    n = p.numel()
                    red = flat[off:off + n].view_as(p.grad).to(p.grad.dtype)
                    if self.average:
                        red = red / self.world
                    p.grad.copy_(red)
                    off += n
            b["arrived"], b["work"], b["flat"], b["members"] = 0, None, None, None

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []
        if self in _active_syncs:
            _active_syncs.remove(self)
