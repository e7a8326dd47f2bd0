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


def padded_pixels(rows, cols, world, chunk_rows=CHUNK_ROWS):
    """Pixels of a frame buffer padded so that every rank owns the same number of whole chunks (>= rows*cols)."""
    return packed_size(rows, cols, world, chunk_rows) * world


def gather_frame(local_full, rows, cols, world, rank, dist=None, chunk_rows=CHUNK_ROWS):
    """local_full: flat int32 torch tensor holding this rank's chunks at their absolute positions (rows*cols entries, or
    padded_pixels(...) entries -- the padded form avoids a copy).  Chunk c lives on rank c % world, so the buffer viewed as
    [chunks_per_rank, world, chunk_pixels] has this rank's share at [:, rank, :]: one strided copy packs it, one all_gather_into_tensor
    moves it, one permuted copy unpacks it.  Returns the assembled rows*cols frame on rank 0 (None elsewhere).
    Works with NCCL (CUDA tensors) and gloo (CPU tensors)."""
    import torch
    if world == 1:
        return local_full[:rows * cols]
    chunk_pix = chunk_rows * cols
    per_rank = packed_size(rows, cols, world, chunk_rows) // chunk_pix
    total = per_rank * world * chunk_pix
    buf = local_full
    if buf.numel() < total:
        buf = torch.zeros(total, dtype=local_full.dtype, device=local_full.device)
        buf[:local_full.numel()] = local_full
    packed = buf[:total].view(per_rank, world, chunk_pix)[:, rank, :].contiguous()
    out = torch.empty((world, per_rank, chunk_pix), dtype=buf.dtype, device=buf.device)
    dist.all_gather_into_tensor(out.view(-1), packed.view(-1))
    if rank != 0:
        return None
    return out.permute(1, 0, 2).reshape(-1)[:rows * cols]


def allgather_records(local, dist, world):
    """All-gather of variable-length record blocks ([n_r, k] tensors, same dtype/device on every rank) in RANK ORDER.
    Counts are exchanged first, blocks are padded to the longest one (all_gather_into_tensor needs equal sizes), then compacted.
    Returns (tensor [sum n_r, k], counts). Works with NCCL (CUDA tensors) and gloo (CPU tensors)."""
    import torch
    if world == 1:
        return local, [int(local.shape[0])]
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    m = max(max(counts), 1)
    k = local.shape[1]
    padded = torch.zeros((m, k), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    out = torch.empty((world * m, k), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * m:r * m + counts[r]] for r in range(world)]).contiguous(), counts


def photon_pass(scene, n_cast, world, rank, dist=None):
    """Photon map of a photon scene on `world` GPUs: rank r emits photon indices photon_range(n_cast, world, r) of every light, the
    canonical-order records are all-gathered over NCCL, and every rank builds the full grid (the reference emits serially,
    myScene.java:961,1009).  Bit-identical to the single-GPU map for any world size.  Returns (stored photons, photon-side kernel launches of this rank)."""
    import torch
    if world == 1:
        st = scene.emit_photons_range(0, n_cast)          # always re-emits; the grid is built by the next render call
        return int(st.photons_stored), int(st.kernel_launches)
    i0, i1 = photon_range(n_cast, world, rank)
    st = scene.emit_photons_range(i0, i1)
    cnt = int(st.photons_stored)
    buf = torch.empty((max(cnt, 1), 6), dtype=torch.float64, device="cuda")
    scene.photons_export_device(buf.data_ptr(), cnt)
    allrec, _ = allgather_records(buf[:cnt], dist, world)
    torch.cuda.current_stream().synchronize()            # the library works on its own stream
    st2 = scene.photons_build_device(allrec.data_ptr(), int(allrec.shape[0]))
    return int(allrec.shape[0]), int(st.kernel_launches) + int(st2.kernel_launches)
