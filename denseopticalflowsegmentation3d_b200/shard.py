"""Frame-pair sharding across ranks (one process per GPU) and the gather of per-frame results.

Pairs of a video are independent (the reference's video loop carries only `prev_frame`, segment.cpp:209-269):
rank r processes a contiguous block of pairs, i.e. frames [first, first + n] inclusive — the boundary frame is
rendered / uploaded by both neighbours.  There is no collective on the per-frame path; after a shard finishes the
per-frame box counts and the packed box records are gathered (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist

from .capi import BOX_DTYPE


def pair_block(rank, world, n_pairs):
    """(first_pair, n_pairs_of_rank): contiguous blocks, the first n_pairs % world ranks get one more."""
    base, extra = divmod(n_pairs, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def frame_range(rank, world, n_pairs):
    """(first_frame, n_frames) the rank needs: n_pairs_of_rank + 1 frames, pair i = frames i, i+1."""
    first, n = pair_block(rank, world, n_pairs)
    return first, (n + 1 if n > 0 else 0)


def gather_boxes(n_boxes, boxes, device=None, group=None):
    """All ranks contribute their per-frame box counts (int32 [n_local]) and packed boxes (BOX_DTYPE [total_local]);
    every rank gets (counts_per_rank: list of int32 arrays, boxes_per_rank: list of BOX_DTYPE arrays) in rank order.
    Variable lengths are exchanged with one all_gather of sizes and one padded all_gather of bytes."""
    world = dist.get_world_size(group)
    device = device or torch.device("cpu")
    n_boxes = np.ascontiguousarray(n_boxes, np.int32)
    boxes = np.ascontiguousarray(boxes, BOX_DTYPE)
    assert int(n_boxes.sum()) == boxes.shape[0]
    sizes = torch.tensor([n_boxes.size, boxes.shape[0]], dtype=torch.int64, device=device)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = [tuple(int(v) for v in s.cpu()) for s in all_sizes]
    max_frames = max(s[0] for s in all_sizes)
    max_boxes = max(s[1] for s in all_sizes)
    cnt = torch.zeros(max(max_frames, 1), dtype=torch.int32, device=device)
    cnt[:n_boxes.size] = torch.from_numpy(n_boxes).to(device)
    raw = torch.zeros(max(max_boxes, 1) * BOX_DTYPE.itemsize, dtype=torch.uint8, device=device)
    if boxes.shape[0]:
        raw[:boxes.nbytes] = torch.from_numpy(boxes.view(np.uint8).reshape(-1).copy()).to(device)
    all_cnt = [torch.zeros_like(cnt) for _ in range(world)]
    all_raw = [torch.zeros_like(raw) for _ in range(world)]
    dist.all_gather(all_cnt, cnt, group=group)
    dist.all_gather(all_raw, raw, group=group)
    counts, out = [], []
    for r, (nf, nb) in enumerate(all_sizes):
        counts.append(all_cnt[r][:nf].cpu().numpy())
        out.append(all_raw[r][:nb * BOX_DTYPE.itemsize].cpu().numpy().view(BOX_DTYPE).copy())
    return counts, out


def _all_gather_flat(out, inp, group):
    """One collective into a preallocated [world, ...] tensor (a single NCCL all-gather on the GPUs)."""
    try:
        dist.all_gather_into_tensor(out, inp, group=group)
    except (RuntimeError, NotImplementedError):  # a backend without the flat form
        dist.all_gather(list(out.unbind(0)), inp, group=group)


def gather_packed(counts, boxes_u8, group=None, trim=False):
    """Gather of per-frame results straight from device buffers (no host round trip): `counts` int32 [C] = boxes held by
    each of this rank's C contexts, `boxes_u8` uint8 [C][cap * 216] = their dense box lists (dofs3d_pack_boxes_dev).
    Two collectives back to back, no host synchronisation in between: the counts, then the boxes at their fixed capacity
    (a few MB per rank: over NVLink / NVSwitch the padding costs less than the host read a trimmed size would need).
    trim=True first reads the largest count of any context (one host sync) and sends only that many boxes per context.
    Returns (all_counts int32 [world][C], all_boxes uint8 [world][C][m * 216]) on the tensors' device."""
    world = dist.get_world_size(group)
    itemsize = BOX_DTYPE.itemsize
    counts = counts.contiguous()
    all_counts = torch.empty((world,) + tuple(counts.shape), dtype=counts.dtype, device=counts.device)
    _all_gather_flat(all_counts, counts, group)
    send = boxes_u8
    if trim:
        m = min(max(int(all_counts.max().item()), 1), boxes_u8.shape[1] // itemsize)
        send = boxes_u8[:, :m * itemsize]
    send = send.contiguous()
    all_boxes = torch.empty((world,) + tuple(send.shape), dtype=send.dtype, device=send.device)
    _all_gather_flat(all_boxes, send, group)
    return all_counts, all_boxes
