"""Multi-GPU test on real devices (skipped on a single-GPU box): the library's own NCCL path -- drt_comm_init + drt_render_distributed -- must
assemble, on rank 0, the same frame a single GPU renders, bit for bit, for plain and photon scenes, equal and ragged chunk shares."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_distributed_render_matches_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 4 if n >= 4 else 2
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1", "--master-port", "29541",
                        os.path.join(ROOT, "tools", "dist_check.py")], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["ok"] and d["world"] == world and len(d["cases"]) == 8, d


def test_world_of_one_needs_no_nccl(drt, gpu_ctx_factory):
    """drt_render_distributed on a context that never joined a communicator is the single-GPU render (and must not touch NCCL)."""
    ctx = gpu_ctx_factory(160, 120)
    s = drt.Scene.from_cli(ctx, "p3_t08.cli")
    a, _ = s.draw()
    b = np.zeros_like(a)
    st = s.draw_distributed(host_ptr=b.ctypes.data, chunk_rows=8)
    assert np.array_equal(a, b) and st.rays_primary == 160 * 120
    ctx.close()
