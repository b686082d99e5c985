"""Row sharding helpers (extension: the reference is single-device). Rows are split on 65,536-row block
boundaries so that block-seeded synthetic data is identical for every world size."""
from __future__ import annotations

GEN_BLOCK_ROWS = 65536


def shard_bounds(n_rows: int, world: int, rank: int, block: int = GEN_BLOCK_ROWS):
  """[lo, hi) of the rows owned by `rank`: contiguous, block-aligned, covering [0, n_rows) exactly."""
  if not (0 <= rank < world):
    raise ValueError("rank %d outside world of %d" % (rank, world))
  nblocks = (n_rows + block - 1) // block
  b0 = nblocks * rank // world
  b1 = nblocks * (rank + 1) // world
  return min(b0 * block, n_rows), min(b1 * block, n_rows)


def broadcast_bytes(payload, src: int = 0, group=None) -> bytes:
  """Every rank returns the bytes `payload` held on rank `src` (used for the 128-byte NCCL unique id)."""
  import torch.distributed as dist
  box = [payload if dist.get_rank(group) == src else None]
  dist.broadcast_object_list(box, src=src, group=group)
  return box[0]


def allgather_bytes(payload: bytes, group=None) -> bytes:
  """Every rank returns the concatenation, in rank order, of the equally long `payload`s of all ranks (used
  for the 64-byte cudaIpc handles of the peer inboxes)."""
  import torch.distributed as dist
  box = [None] * dist.get_world_size(group)
  dist.all_gather_object(box, payload, group=group)
  if len(set(len(b) for b in box)) != 1:
    raise ValueError("allgather_bytes: payload lengths differ across ranks")
  return b"".join(box)
