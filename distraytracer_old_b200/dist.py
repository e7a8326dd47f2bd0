"""Multi-GPU host logic: one process per GPU, image partitioned by interleaved row chunks, photons by index range.

The reference is single threaded (myScene.java:15 declares an unused executor); pixels are independent
(myScene.java:1499-1519), so the path shards with no data-path collective.  The only exchanges are
  * gather of the finished framebuffer chunks to rank 0 (all_gather of equal-size packed chunk buffers), and
  * all-gather of photon records when emission is split per rank (see photon path).
Results are bit-identical for any world size because the sampler is keyed by absolute pixel / sample index
and a pixel's samples are never split across ranks (the per-pixel mean is a sequential sum).
"""
import numpy as np

CHUNK_ROWS = 8


def chunk_table(rows, cols, world, chunk_rows=CHUNK_ROWS):
    """Return per-rank list of (pix0, pix1) pixel ranges (row-major), round-robin over row chunks."""
    n_chunks = (rows + chunk_rows - 1) // chunk_rows
    table = [[] for _ in range(world)]
    for c in range(n_chunks):
        r0, r1 = c * chunk_rows, min(rows, (c + 1) * chunk_rows)
        table[c % world].append((r0 * cols, r1 * cols))
    return table


def packed_size(rows, cols, world, chunk_rows=CHUNK_ROWS):
    """Pixels of the per-rank packed buffer (equal on every rank: padded to the largest share)."""
    n_chunks = (rows + chunk_rows - 1) // chunk_rows
    per_rank = (n_chunks + world - 1) // world
    return per_rank * chunk_rows * cols


def pack_index(rows, cols, world, rank, chunk_rows=CHUNK_ROWS):
    """int64 array: absolute pixel index for every slot of rank's packed buffer (-1 = padding)."""
    idx = np.full(packed_size(rows, cols, world, chunk_rows), -1, dtype=np.int64)
    at = 0
    for p0, p1 in chunk_table(rows, cols, world, chunk_rows)[rank]:
        idx[at:at + (p1 - p0)] = np.arange(p0, p1)
        at += chunk_rows * cols
    return idx


def photon_range(n_cast, world, rank):
    """Rank r emits global photon indices [r*N/n, (r+1)*N/n) of every light."""
    return (n_cast * rank) // world, (n_cast * (rank + 1)) // world


def gather_frame(local_full, rows, cols, world, rank, dist=None, chunk_rows=CHUNK_ROWS):
    """local_full: flat int32 torch tensor (rows*cols) holding this rank's chunks at their absolute positions.
    Returns the assembled frame on rank 0 (None elsewhere).  Works with NCCL (CUDA tensors) and gloo (CPU tensors)."""
    import torch
    if world == 1:
        return local_full
    idx = torch.from_numpy(pack_index(rows, cols, world, rank, chunk_rows)).to(local_full.device)
    valid = idx >= 0
    packed = torch.zeros(idx.numel(), dtype=local_full.dtype, device=local_full.device)
    packed[valid] = local_full[idx[valid]]
    out = torch.empty(world * packed.numel(), dtype=local_full.dtype, device=local_full.device)
    dist.all_gather_into_tensor(out, packed)
    if rank != 0:
        return None
    frame = torch.zeros(rows * cols, dtype=local_full.dtype, device=local_full.device)
    for r in range(world):
        ridx = torch.from_numpy(pack_index(rows, cols, world, r, chunk_rows)).to(local_full.device)
        v = ridx >= 0
        frame[ridx[v]] = out[r * packed.numel():(r + 1) * packed.numel()][v]
    return frame
