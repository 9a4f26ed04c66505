"""Shard-level multi-GPU plumbing.  Shard proofs are independent in this reference version
(every `open` starts from a clone of the same challenger, crates/stark/src/prover.rs:687 and
crates/core/machine/src/utils/prove.rs:496), so the path shards with NO data-path collective:
rank r proves shards r, r+world, ...  The only exchange is the gather of the 8-word commitments /
proof blobs to the rank that assembles `MachineProof{shard_proofs}` (prove.rs:570)."""
from __future__ import annotations

import numpy as np


def assign_shards(n_shards: int, world: int, rank: int) -> list[int]:
    """Round-robin partition (the order `records.into_par_iter()` would hand them out)."""
    return list(range(rank, n_shards, world))


def gather_commitments(local: dict[int, np.ndarray], n_shards: int, group=None, device="cpu") -> np.ndarray | None:
    """All ranks contribute {shard index: commitment[8]}; every rank gets the (n_shards, 8) table.
    Works over NCCL (device='cuda') and gloo (device='cpu')."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    per = -(-n_shards // world)
    buf = torch.full((per, 9), -1, dtype=torch.int64, device=device)
    for k, (idx, c) in enumerate(sorted(local.items())):
        buf[k, 0] = idx
        buf[k, 1:] = torch.from_numpy(np.asarray(c, dtype=np.int64)).to(device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    table = np.zeros((n_shards, 8), dtype=np.uint32)
    seen = set()
    for t in out:
        for row in t.cpu().numpy():
            if row[0] >= 0:
                table[int(row[0])] = row[1:].astype(np.uint32)
                seen.add(int(row[0]))
    assert seen == set(range(n_shards)), "a shard commitment is missing"
    return table
