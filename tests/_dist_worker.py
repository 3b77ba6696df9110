"""Worker for the multi-process (gloo) tests: one process per rank, exactly as bench.py runs one
process per GPU.  Writes its shared-node lists to `outdir/rank{r}.npz`."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def worker(rank, nranks, port, ne, lx, outdir):
    import torch
    import torch.distributed as dist
    import neko_top_b200  # noqa: F401
    from neko_top_b200 import partition, workloads
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=nranks)
    try:
        brick = workloads.config_weak(rank, nranks, ne_per_gpu=ne, lx=lx)
        keys = workloads.node_keys(brick).reshape(-1)
        cand = workloads.interface_candidates(brick)
        sh = partition.find_shared_nodes(keys, cand, lx ** 3, rank, nranks)
        np.savez(os.path.join(outdir, f"rank{rank}.npz"), keys=keys.numpy(), shared_key=sh.shared_key,
                 shared_dof=sh.shared_dof, neigh_rank=sh.neigh_rank, neigh_off=sh.neigh_off,
                 neigh_idx=sh.neigh_idx, bnd_elem=sh.bnd_elem)
        dist.barrier()
    finally:
        dist.destroy_process_group()
