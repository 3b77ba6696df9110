"""Multi-GPU parity check (run under torch.distributed.run, one process per GPU):
every rank computes the fused step on its brick with the NCCL shared-node exchange, rank-local results
are compared with the CPU oracle evaluated on the UNDIVIDED global mesh (oracle = checker only).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tools/mgpu_check.py [ne_per_gpu] [lx]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import neko_top_b200  # noqa: E402,F401
from neko_top_b200 import operators as ops, partition, sem, workloads  # noqa: E402


def main():
    ne = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    lx = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sp = sem.Space(lx)

    def build(brick, device):
        x, y, z = workloads.coords(brick, device)
        keys = workloads.node_keys(brick, device)
        G, _, B = sem.geometric_factors(x, y, z, sp)
        fl = workloads.make_fields(brick, x, y, z, keys)
        flat = lambda a: a.reshape(-1).contiguous()
        return dict(G=[flat(g) for g in G], B=flat(B), v=[flat(a) for a in fl.v], ub=[flat(a) for a in fl.ub],
                    rho=flat(fl.rho), keys=flat(keys))

    brick = workloads.config_weak(rank, world, ne, lx)
    brick.deform = 0.02
    d = build(brick, dev)
    op = ops.fused_adjoint_rhs_t(ops.coef_t(ops.space_t(lx, sp.dx, sp.wx), brick.nelv, d["G"], d["B"]), device=local)
    op.gs.init(d["keys"])
    idb = [ops.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(idb, src=0)
    op.comm_init(idb[0], rank, world)
    # discovery behind the C ABI (device + ncclAllGather), cross-checked against the torch.distributed host
    # implementation the CPU (gloo) tests cover
    cand = workloads.interface_candidates(brick, dev)
    sh = partition.find_shared_nodes(d["keys"], cand, lx ** 3, rank, world)
    ns_abi, nn_abi = op.gs.init_shared_from_keys(d["keys"], cand)
    assert ns_abi == sh.nshared and nn_abi == sh.neigh_rank.size, (ns_abi, sh.nshared, nn_abi, sh.neigh_rank.size)
    if world == 2:        # no candidate mask: every element-surface dof is a candidate
        ns2, nn2 = op.gs.init_shared_from_keys(d["keys"].cpu().numpy(), None)
        assert (ns2, nn2) == (ns_abi, nn_abi)
    n = brick.n
    mk = lambda: [torch.full((n,), float("nan"), device=dev, dtype=torch.float64) for _ in range(3)]
    results = {}
    for mode in ("staged", "sequential", "no_xstage", "overlap", "overlap_gs1"):
        # sequential = the default path (exchange overlapped with the local gs); with
        # B200_EXCHANGE_OVERLAP=elem the "overlap*" modes run the boundary/interior split
        if mode == "staged":                   # the default: product classes summed direction by direction
            op.set_xstage(2)
        if mode == "sequential":               # x pairs in the kernel only: bit-identical to every mode below
            op.set_xstage(1)
        if mode == "no_xstage":                # plain element kernel + full local pass
            op.set_xstage(0)
        if mode == "overlap":
            op.set_xstage(1)
            op.set_boundary_elements(sh.bnd_elem)
        if mode.startswith("overlap_gs"):      # packed class lists
            op.set_gs_mode(int(mode[-1]))
        f, sens = mk(), torch.empty(n, device=dev, dtype=torch.float64)
        for _ in range(2):
            op.step(d["v"], d["ub"], f, rho=d["rho"], sens=sens)
        torch.cuda.synchronize()
        if mode == "staged":
            xs_info = op.xstage_info()
        results[mode] = [a.cpu().numpy() for a in f] + [sens.cpu().numpy()]
    # host-buffer path
    hv = [a.cpu().pin_memory() for a in d["v"]]
    hub = [a.cpu().pin_memory() for a in d["ub"]]
    hf = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(3)]
    hs = torch.empty(n, dtype=torch.float64).pin_memory()
    op.set_gs_mode(0)
    op.step_host(hv, hub, d["rho"].cpu().pin_memory(), hf, hs)
    results["host"] = [a.numpy() for a in hf] + [hs.numpy()]

    # oracle on the undivided mesh (CPU; small sizes only)
    from oracle import pyoracle as orc
    px, py, pz = workloads.rank_grid(world)
    whole = workloads.BoxBrick(lx=lx, ne=(ne * px, ne * py, ne * pz), length=(float(px), float(py), float(pz)),
                               deform=0.02)
    w = build(whole, "cpu")
    c = lambda t: t.numpy()
    fo, so, _ = orc.adjoint_rhs([c(a) for a in w["v"]], [c(a) for a in w["ub"]], lx, whole.nelv, sp.dx, sp.wx,
                                [c(g) for g in w["G"]], c(w["B"]), rho=c(w["rho"]))
    cid, nc = orc.gs_classes(c(w["keys"]))
    fo = [orc.gs_add(a, cid, nc) for a in fo]
    e = np.arange(brick.nelv)
    ex, ey, ez = e % ne + brick.offset[0], (e // ne) % ne + brick.offset[1], e // (ne * ne) + brick.offset[2]
    ge = ex + whole.ne[0] * (ey + whole.ne[1] * ez)
    N = lx ** 3
    ref = [a.reshape(whole.nelv, N)[ge].reshape(-1) for a in fo] + [so.reshape(whole.nelv, N)[ge].reshape(-1)]
    worst = 0.0
    for mode, res in results.items():
        for a, b in zip(res, ref):
            err = np.linalg.norm(a - b) / np.linalg.norm(b)
            worst = max(worst, err)
        print(f"rank {rank} {mode}: max rel-L2 vs global oracle = {worst:.3e}", flush=True)
    op.set_gs_mode(0)
    same = all(np.array_equal(a, b) for m in ("no_xstage", "overlap", "overlap_gs1")
               for a, b in zip(results["sequential"], results[m]))
    # the staged sums associate 4- and 8-member classes pairwise: equal to the class-list sums up to rounding
    same = same and all(np.abs(a - b).max() <= 4e-15 * np.abs(b).max()
                        for a, b in zip(results["staged"], results["sequential"]))
    same_h = all(np.array_equal(a, b) for a, b in zip(results["sequential"], results["host"]))
    t = torch.tensor([worst, 0.0 if (same and same_h) else 1.0], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = True
    if rank == 0:
        ok = t[0].item() <= 1e-12 and t[1].item() == 0.0
        print(f"MGPU_CHECK world={world} ne={ne} lx={lx} nshared(rank0)={sh.nshared} nbnd(rank0)={sh.bnd_elem.size} "
              f"xstage(rank0: level, classes staged, classes left, total)={xs_info} "
              f"max_err={t[0].item():.3e} modes_bit_identical={t[1].item() == 0.0} -> {'PASS' if ok else 'FAIL'}",
              flush=True)
    op.free()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
