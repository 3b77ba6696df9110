"""Developer tool: step time with the direct-stiffness summation inside the element kernel vs the
separate pass, for several element processing orders.  Usage: python tools/gsbench.py [ne] [tiles...]
tiles: e.g. mesh 8x8 16x16 32x8 (mesh = lexicographic order)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import neko_top_b200  # noqa: E402,F401
from neko_top_b200 import operators as ops, sem, workloads  # noqa: E402


def main():
    ne = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    tiles = sys.argv[2:] if len(sys.argv) > 2 else ["sep", "fused"]
    lx, dev = 8, "cuda"
    brick = workloads.config_box(ne, lx)
    sp = sem.Space(lx)
    x, y, z = workloads.coords(brick, dev)
    keys = workloads.node_keys(brick, dev)
    G, jac, B = sem.geometric_factors(x, y, z, sp, chunk=8192)
    fl = workloads.make_fields(brick, x, y, z, keys)
    del x, y, z, jac
    flat = lambda a: a.reshape(-1).contiguous()
    G = [flat(g) for g in G]
    B = flat(B)
    v, ub, rho = [flat(a) for a in fl.v], [flat(a) for a in fl.ub], flat(fl.rho)
    del fl
    n = brick.n
    f = [torch.empty(n, device=dev, dtype=torch.float64) for _ in range(3)]
    sens = torch.empty(n, device=dev, dtype=torch.float64)
    coef = ops.coef_t(ops.space_t(lx, sp.dx, sp.wx), brick.nelv, G, B)
    kflat = keys.reshape(-1)
    ref = None
    # spec: csr | sep | fused, then [:lag=N][:hint=0/1][:tile=AxB][:cfg=N]
    for spec in tiles:
        kv = dict(a.split("=") for a in spec.split(":")[1:])
        os.environ["B200_GS_LAG"] = kv.get("lag", "2")
        os.environ["B200_GS_L2HINT"] = kv.get("hint", "1")
        os.environ["B200_GS_UN"] = kv.get("un", "1")
        os.environ["B200_ADJRHS_CFG"] = kv.get("cfg", "-1")
        op = ops.fused_adjoint_rhs_t(coef)
        op.gs.init(kflat)
        op.set_gs_mode(2 if spec.startswith("fused") else (0 if spec.startswith("csr") else 1))
        if "tile" in kv:
            tx, ty = (int(a) for a in kv["tile"].split("x"))
            op.set_element_order(workloads.tile_order(brick, (tx, ty)))
        for _ in range(3):
            op.step(v, ub, f, rho=rho, sens=sens)
        torch.cuda.synchronize()
        op.enable_timing(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            op.step(v, ub, f, rho=rho, sens=sens)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        ek, gk, _ = op.get_timing()
        same = None
        if ref is None:
            ref = f[0].clone()
        else:
            same = bool(torch.equal(f[0], ref))
        print(f"{spec:36s} step {ms:.4f} ms = {n / ms / 1e6:.2f} GDOF/s ({200.4 * n / ms / 1e6:.0f} GB/s alg.) "
              f"elem {ek:.4f} gs {gk:.4f}  fused={op.gs_info()[0]} same={same}", flush=True)
        op.free()


if __name__ == "__main__":
    main()
