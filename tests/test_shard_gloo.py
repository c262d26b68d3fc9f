"""CPU tests of the multi-GPU host logic with world_size 2 over gloo: contiguous frame-pair blocks and the gather of
variable-length per-frame box lists (what runs over NCCL/NVLink on the GPUs)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from denseopticalflowsegmentation3d_b200 import shard
from denseopticalflowsegmentation3d_b200.capi import BOX_DTYPE


def test_pair_blocks_partition_the_stream():
    for n in (0, 1, 7, 512, 8192):
        for world in (1, 2, 3, 8):
            blocks = [shard.pair_block(r, world, n) for r in range(world)]
            assert blocks[0][0] == 0 and sum(b[1] for b in blocks) == n
            for a, b in zip(blocks, blocks[1:]):
                assert a[0] + a[1] == b[0]
            assert max(b[1] for b in blocks) - min(b[1] for b in blocks) <= 1
            for r in range(world):
                first, nf = shard.frame_range(r, world, n)
                assert nf == (blocks[r][1] + 1 if blocks[r][1] else 0) and first == blocks[r][0]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_pairs, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, n = shard.pair_block(rank, world, n_pairs)
    rng = np.random.default_rng(100 + rank)
    n_boxes = rng.integers(0, 4, size=n).astype(np.int32)
    if rank == 1:
        n_boxes[:] = 0  # a rank without any box
    boxes = np.zeros(int(n_boxes.sum()), BOX_DTYPE)
    boxes["root"] = first * 1000 + np.arange(boxes.shape[0])
    boxes["score"] = rank + 0.5
    counts, gathered = shard.gather_boxes(n_boxes, boxes)
    # the fixed-capacity form used on the GPUs (tensors stay where they are): two "contexts" per rank
    import torch
    cap = 16
    packed = torch.zeros((2, cap * BOX_DTYPE.itemsize), dtype=torch.uint8)
    half = boxes.shape[0] // 2
    parts = [boxes[:half], boxes[half:]]
    for c, part in enumerate(parts):
        packed[c, :part.nbytes] = torch.from_numpy(part.view(np.uint8).reshape(-1).copy())
    cnt = torch.tensor([p.shape[0] for p in parts], dtype=torch.int32)
    all_cnt, all_box = shard.gather_packed(cnt, packed, trim=(rank + n_pairs) % 2 == 0 or True)
    full_cnt, full_box = shard.gather_packed(cnt, packed)
    assert torch.equal(full_cnt, all_cnt) and full_box.shape[2] == cap * BOX_DTYPE.itemsize
    for r in range(world):
        for c in range(2):
            k = int(all_cnt[r, c])
            got = all_box[r, c, :k * BOX_DTYPE.itemsize].numpy().view(BOX_DTYPE)
            assert np.all(got["score"] == r + 0.5)
            if r == rank:
                assert got["root"].tolist() == parts[c]["root"].tolist()
    assert int(all_cnt.sum()) == sum(int(c.sum()) for c in counts)
    out_q.put((rank, [c.tolist() for c in counts], [g["root"].tolist() for g in gathered], [g["score"].tolist() for g in gathered],
               n_boxes.tolist(), boxes["root"].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_boxes_world2_gloo():
    world, n_pairs = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # every rank sees every rank's contribution, in rank order
    own = {r: (res[r][4], res[r][5]) for r in range(world)}
    for rank, counts, roots, scores, _, _ in res:
        for r in range(world):
            assert counts[r] == own[r][0] and roots[r] == own[r][1]
            assert all(s == r + 0.5 for s in scores[r])
    assert len(res[0][1][0]) == 4 and len(res[0][1][1]) == 3  # 7 pairs -> blocks of 4 and 3
