"""world_size-2 gloo test (CPU) of the multi-GPU host logic: chunk partition + frame gather."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    from distraytracer_old_b200 import dist as D
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    rows, cols = 37, 29
    # stand-in renderer: every rank fills only its own chunks with f(pixel); everything else stays 0
    local = torch.zeros(rows * cols, dtype=torch.int32)
    for p0, p1 in D.chunk_table(rows, cols, world)[rank]:
        local[p0:p1] = torch.arange(p0, p1, dtype=torch.int32) * 7 + 1
    frame = D.gather_frame(local, rows, cols, world, rank, dist)
    if rank == 0:
        want = torch.arange(rows * cols, dtype=torch.int32) * 7 + 1
        assert torch.equal(frame, want), "assembled frame differs"
        print("GATHER_OK")
    else:
        assert frame is None
    # same through a buffer padded to whole chunks per rank (what bench.py allocates: the pack is then a strided view, no copy)
    padded = torch.zeros(D.padded_pixels(rows, cols, world), dtype=torch.int32)
    padded[:rows * cols] = local
    frame2 = D.gather_frame(padded, rows, cols, world, rank, dist)
    if rank == 0:
        assert torch.equal(frame2, want), "assembled frame (padded buffer) differs"
    # photon exchange: ranks own contiguous photon-index ranges with a different number of stored records each; the gathered set must be
    # the rank-order concatenation (== the single-GPU canonical order)
    n_cast = 1003
    i0, i1 = D.photon_range(n_cast, world, rank)
    mine = torch.stack([torch.arange(i0, i1, dtype=torch.float64) * 10 + k for k in range(6)], dim=1)[::(2 + rank)].contiguous()   # ragged counts
    allrec, counts = D.allgather_records(mine, dist, world)
    want = torch.cat([torch.stack([torch.arange(*D.photon_range(n_cast, world, r), dtype=torch.float64) * 10 + k for k in range(6)], dim=1)[::(2 + r)] for r in range(world)])
    assert counts == [len(range(*D.photon_range(n_cast, world, r))[::(2 + r)]) for r in range(world)]
    assert torch.equal(allrec, want), "gathered photon records differ"
    empty, c0 = D.allgather_records(mine[:0], dist, world)          # nobody stored anything
    assert empty.shape[0] == 0 and c0 == [0] * world
    if rank == 0:
        print("PHOTONS_OK")
    # ---- the partition the LIBRARY uses (drt_render_distributed; pure host functions behind the C ABI): rank r fills its compact buffer, rank 0
    #      receives every rank's buffer (what ncclSend / ncclRecv do on the GPUs) and unpacks it with the same mapping -> the identity frame
    from distraytracer_old_b200 import host as H
    for rows2, cols2, cr in ((37, 29, 8), (40, 16, 5), (7, 3, 8)):
        idx = torch.from_numpy(H.dist_abs_pixels(cols2, rows2, world, rank, cr))
        assert idx.numel() == H.dist_rank_pixels(cols2, rows2, world, rank, cr)
        mine = torch.where(idx >= 0, idx * 3 + 5, torch.full_like(idx, -99))
        sizes = [H.dist_rank_pixels(cols2, rows2, world, r, cr) for r in range(world)]
        bufs = [torch.zeros(sizes[r], dtype=torch.int64) for r in range(world)] if rank == 0 else None
        if rank == 0:
            bufs[0] = mine
            for r in range(1, world):
                dist.recv(bufs[r], src=r)
            frame = torch.full((rows2 * cols2,), -1, dtype=torch.int64)
            for r in range(world):
                ir = torch.from_numpy(H.dist_abs_pixels(cols2, rows2, world, r, cr)); ok = ir >= 0
                frame[ir[ok]] = bufs[r][ok]
            assert torch.equal(frame, torch.arange(rows2 * cols2, dtype=torch.int64) * 3 + 5), "library partition does not tile the frame"
        else:
            dist.send(mine, dst=0)
    assert H.dist_photon_range(1003, world, rank) == D.photon_range(1003, world, rank)
    i01 = [H.dist_photon_range(1003, world, r) for r in range(world)]
    assert i01[0][0] == 0 and i01[-1][1] == 1003 and all(i01[r][1] == i01[r + 1][0] for r in range(world - 1))
    if rank == 0:
        print("LIBPART_OK")
    dist.barrier(); dist.destroy_process_group()
""") % ROOT


def test_frame_gather_world2(tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", str(w)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GATHER_OK" in r.stdout and "PHOTONS_OK" in r.stdout and "LIBPART_OK" in r.stdout
