"""Pose-batch sharding across the GPUs of one box.

The path shards trivially: queries are independent and the models are read-only, so BVHs are
replicated on every GPU, the pose batch is split into contiguous blocks (rank r owns
[r*n/G, (r+1)*n/G)), and the per-rank result records are gathered with one all_gather per
output array (NCCL over NVLink on GPUs; the same code runs over gloo on CPU tensors, which is
how the host-side logic is tested without GPUs).  There is no exchange step inside the
traversal, hence no other collective.
"""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous block of rank `rank`: balanced to within one query."""
    base, rem = divmod(int(n), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_sizes(n, world):
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def all_gather_records(local, n_total, group=None):
    """Gathers per-rank result tensors (first dim = queries of that rank, in rank order) into one
    tensor of n_total rows on every rank.  Ragged shards (n_total % world != 0) are padded to the
    largest shard for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = shard_sizes(n_total, world)
    if world == 1:
        return local
    mx = max(sizes)
    if local.shape[0] != mx:
        pad = torch.zeros((mx - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    if all(s == mx for s in sizes):
        return out
    parts = [out[r * mx: r * mx + sizes[r]] for r in range(world)]
    return torch.cat(parts, dim=0)


def all_gather_contacts(num_contacts, contacts_u8, n_total, group=None):
    """Ragged contact lists: gather the counts, then the (padded) 64-byte contact records.
    Returns (counts[n_total], offsets[n_total+1], contacts_u8[total*64]) on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    counts = all_gather_records(num_contacts, n_total, group)
    offsets = torch.zeros(n_total + 1, dtype=torch.int64, device=counts.device)
    torch.cumsum(counts.to(torch.int64), dim=0, out=offsets[1:])
    if world == 1:
        return counts, offsets, contacts_u8
    my_bytes = torch.tensor([contacts_u8.numel()], dtype=torch.int64, device=contacts_u8.device)
    all_bytes = torch.empty(world, dtype=torch.int64, device=contacts_u8.device)
    dist.all_gather_into_tensor(all_bytes, my_bytes, group=group)
    sizes = [int(x) for x in all_bytes.tolist()]
    mx = max(max(sizes), 64)
    buf = torch.zeros(mx, dtype=torch.uint8, device=contacts_u8.device)
    buf[: contacts_u8.numel()] = contacts_u8.reshape(-1)
    out = torch.empty(world * mx, dtype=torch.uint8, device=contacts_u8.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    parts = [out[r * mx: r * mx + sizes[r]] for r in range(world)]
    return counts, offsets, torch.cat(parts)


def numpy_shard(arr, rank, world):
    s, e = shard_range(len(arr), rank, world)
    return np.ascontiguousarray(arr[s:e])
