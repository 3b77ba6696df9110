"""Developer tool (not the contract bench): time the fused element kernel and the gs pass for each
kernel configuration (B200_ADJRHS_CFG) on a box mesh.  Usage: python tools/kbench.py [ne] [lx] [cfgs]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import neko_top_b200  # noqa: E402,F401
from neko_top_b200 import operators as ops, sem, workloads  # noqa: E402


def main():
    ne = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    lx = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    cfgs = [int(c) for c in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
    dev = "cuda"
    brick = workloads.config_box(ne, lx)
    sp = sem.Space(lx)
    x, y, z = workloads.coords(brick, dev)
    keys = workloads.node_keys(brick, dev)
    G, jac, B = sem.geometric_factors(x, y, z, sp)
    fl = workloads.make_fields(brick, x, y, z, keys)
    del x, y, z, jac
    flat = lambda a: a.reshape(-1).contiguous()
    G = [flat(g) for g in G]
    B = flat(B)
    v, ub, rho = [flat(a) for a in fl.v], [flat(a) for a in fl.ub], flat(fl.rho)
    n = brick.n
    f = [torch.empty(n, device=dev, dtype=torch.float64) for _ in range(3)]
    sens = torch.empty(n, device=dev, dtype=torch.float64)
    bpd = sem.algorithmic_bytes_per_dof(lx, with_gs=False)
    for cfg in cfgs:
        os.environ["B200_ADJRHS_CFG"] = str(cfg)
        coef = ops.coef_t(ops.space_t(lx, sp.dx, sp.wx), brick.nelv, G, B)
        op = ops.fused_adjoint_rhs_t(coef)
        t0 = time.time()
        op.gs.init(keys.reshape(-1))
        torch.cuda.synchronize()
        t_gs_init = time.time() - t0
        for _ in range(3):
            op.step(v, ub, f, rho=rho, sens=sens)
        torch.cuda.synchronize()
        op.enable_timing(True)
        for _ in range(10):
            op.step(v, ub, f, rho=rho, sens=sens)
        ek, gk, nl = op.get_timing()
        op.enable_timing(False)
        print(f"cfg={cfg} lx={lx} ne={ne}^3 n={n}: elem {ek:.4f} ms = {n / ek / 1e6:.2f} GDOF/s = "
              f"{n * bpd / ek / 1e6:.0f} GB/s ; gs {gk:.4f} ms ; step {n / (ek + gk) / 1e6:.2f} GDOF/s "
              f"; gs_init {t_gs_init:.2f} s", flush=True)
        op.free()


if __name__ == "__main__":
    main()
