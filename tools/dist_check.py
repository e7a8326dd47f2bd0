#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun (one process per GPU):  torchrun --nproc-per-node N tools/dist_check.py
Every rank renders its interleaved row chunks through drt_render_distributed (the library's own NCCL communicator: chunk gather by send/recv,
photon records by all-gather); rank 0 then renders the same frames alone and compares them bit for bit.  Prints one JSON line."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import distraytracer_old_b200 as drt

CASES = [("p3_t09.cli", 640, 360, 2, -1), ("planets3columns.cli", 333, 250, 2, -1), ("box_caustics.cli", 320, 240, 1, 300000), ("t11.cli", 200, 150, 2, 40000)]

def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(drt.Context.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    out = {"world": world, "cases": []}
    for name, cols, rows, spp, photons in CASES:
        ctx = drt.Context(device=local, cols=cols, rows=rows)
        ctx.comm_init(bytes(idt.cpu().numpy().tobytes()), world, rank)      # one id per communicator: re-made per case below
        s = drt.Scene.from_cli(ctx, name, spp=spp, photons=photons, accel=drt.ACCEL_REFERENCE_FAST)
        host = np.zeros((rows, cols), dtype=np.int32)
        for chunk_rows in (8, 5):                                            # 5: ragged last chunk, unequal shares
            st = s.draw_distributed(host_ptr=host.ctypes.data if rank == 0 else None, chunk_rows=chunk_rows, reemit_photons=True)
            if rank == 0:
                solo_ctx = drt.Context(device=local, cols=cols, rows=rows)
                ref, st1 = drt.Scene.from_cli(solo_ctx, name, spp=spp, photons=photons, accel=drt.ACCEL_REFERENCE_FAST).draw()
                solo_ctx.close()
                out["cases"].append({"scene": name, "chunk_rows": chunk_rows, "identical": bool(np.array_equal(ref, host)), "photons_stored": int(st.photons_stored), "photons_stored_single": int(st1.photons_stored)})
            dist.barrier()
        ctx.close()
        # a fresh id for the next context's communicator
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(drt.Context.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
    if rank == 0:
        out["ok"] = all(c["identical"] and c["photons_stored"] == c["photons_stored_single"] for c in out["cases"])
        print(json.dumps(out), flush=True)
    dist.barrier(); dist.destroy_process_group()

if __name__ == "__main__":
    main()
