"""Multi-GPU plumbing of the hot path: rays are independent, so ranks take
contiguous ray ranges *within each image* (the per-image grouping of get_feat_loss_corr, loss.py:120-123,
stays local), weights and feature maps are replicated, and the only exchange is ONE all-reduce(SUM) of the
loss partials -- rgb (sum, n_rays) and per-image feature (sum, count) -- after which every rank forms the
scalars exactly like loss.py:27 and :155-163.  torch.distributed (NCCL on GPUs, gloo in the CPU tests)
is only the transport."""
from __future__ import annotations

from typing import Dict

import torch
import torch.distributed as dist

RAY_KEYS = ("uv", "object_mask", "rgb")


def shard_bounds(n_pixels: int, rank: int, world: int):
    """[begin, end) of the rank's contiguous slice of one image's rays (remainder spread over the first ranks)."""
    base, rem = divmod(n_pixels, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_index(n_pixels: int, rank: int, world: int, tile: int = 0) -> torch.Tensor:
    """Ray indices (within one image) owned by `rank`.  tile = 0: one contiguous range (shard_bounds).  tile > 0: ray tiles
    of `tile` consecutive rays dealt round-robin (tile j goes to rank j % world) -- SURVEY.md section 8e's remedy for load
    imbalance: the hit fraction and the sphere-tracing iteration count vary across the image (the object sits in the
    middle rows), so contiguous ranges give the outer ranks mostly sphere-miss rays."""
    if tile <= 0:
        b, e = shard_bounds(n_pixels, rank, world)
        return torch.arange(b, e)
    idx = torch.arange(n_pixels)
    return idx[(idx // tile) % world == rank]


def shard_rays(batch: Dict[str, torch.Tensor], rank: int, world: int, tile: int = 0) -> Dict[str, torch.Tensor]:
    """Per-image ray slices for this rank (contiguous, or round-robin tiles when tile > 0); everything that is not
    per-ray is replicated."""
    n = batch["uv"].shape[1]
    out = dict(batch)
    if tile <= 0:
        b, e = shard_bounds(n, rank, world)
        for k in RAY_KEYS:
            if k in batch:
                out[k] = batch[k][:, b:e].contiguous()
        return out
    idx = shard_index(n, rank, world, tile)
    for k in RAY_KEYS:
        if k in batch:
            out[k] = batch[k][:, idx].contiguous()
    return out


def mean_from_partials(local_sum: torch.Tensor, count: int, reduce_fn=None) -> torch.Tensor:
    """Mean of a per-sample term whose samples are spread over the ranks: (sum over ALL ranks) / (count over ALL ranks).
    The VALUE is the global mean on every rank; the GRADIENT is d(local_sum) / global_count, so that summing the parameter
    gradients over the ranks (allreduce_gradients) gives exactly the single-GPU gradient.  reduce_fn = None: plain local
    mean.  Used for the eikonal and surface-indicator terms (loss.py:30-35, :167-174), which the reference forms with
    .mean() over tensors that are sharded here."""
    if reduce_fn is None:
        return local_sum / max(count, 1)
    part = torch.stack([local_sum.detach().to(torch.float64),
                        torch.tensor(float(count), dtype=torch.float64, device=local_sum.device)])
    reduce_fn(part)
    tot, cnt = part[0].to(local_sum.dtype), part[1].clamp_min(1.0).to(local_sum.dtype)
    return (local_sum - local_sum.detach() + tot) / cnt


def allreduce_partials(*partials: torch.Tensor, group=None) -> None:
    """The single collective of the path: SUM over ranks, in place, one flat buffer."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([p.reshape(-1) for p in partials])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for p in partials:
        p.copy_(flat[off:off + p.numel()].view_as(p))
        off += p.numel()


def finalize_rgb(partial: torch.Tensor) -> torch.Tensor:
    """partial = (sum |rgb-gt| over hits, n_rays)  ->  loss.py:27."""
    return (partial[0] / partial[1]).to(torch.float32) if float(partial[1]) > 0 else torch.zeros((), dtype=torch.float32)


def finalize_feat(partials: torch.Tensor) -> torch.Tensor:
    """partials [B,2] = (sum of kept |1-corr|, (V-1) m_i)  ->  mean over images of per-image means (loss.py:155-163)."""
    s, c = partials[:, 0], partials[:, 1]
    per_image = torch.where(c > 0, s / c.clamp_min(1), torch.zeros_like(s)).to(torch.float32)
    return per_image.sum() / partials.shape[0]


def allreduce_gradients(params, world_size: int = None, group=None) -> None:
    """Data-parallel training (outside the north-star forward metric): after ``loss.backward()`` every
    rank holds the gradient of ITS rays' share of the loss.  With the loss partials all-reduced before the scalars are
    formed (``B200IDRLoss.forward(reduce_fn=...)``: rgb, feature, depth AND -- through mean_from_partials -- the eikonal and
    surface-indicator means) each rank's backward already uses the GLOBAL denominators, so the global
    gradient is the SUM over ranks -- one all-reduce over a single flat fp32 bucket (2.9 M floats for the 8x512 + 4x512
    networks: latency-bound over NVLink, no bucketing needed).  Parameters without a gradient contribute zeros so that
    all ranks build the same bucket."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    params = [p for p in params if p.requires_grad]
    if not params:
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(torch.float32) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
