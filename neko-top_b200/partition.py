"""Shared-node discovery between ranks: the host-side set-up that Neko's gs_t%init performs from the
dofmap (/root/reference/sources/adjoint/adjoint_scheme.f90:339-343 `gs_Xh%init(dm_Xh)`; the exchange
it prepares is the one behind gs_Xh%op at adjoint/adjoint_pnpn.f90:725,755-757).

One process per GPU.  Every rank hands in the global node key of each local dof and a
boolean mask of the dofs that can possibly live on another rank (partition-boundary candidates).  The
ranks all-gather the sorted unique candidate keys (torch.distributed: NCCL on the GPU box, gloo in the
CPU tests); the intersection of two ranks' key sets, in ascending key order, is the per-neighbour
message layout both sides agree on without any further handshake.

Returns exactly what include/neko_top_b200.h:b200_gs_init_shared and
b200_adjrhs_set_boundary_elements take.
"""
from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist


@dataclass
class SharedNodes:
    shared_key: np.ndarray    # int64 [nshared] ascending global keys of nodes shared with >= 1 other rank
    shared_dof: np.ndarray    # int32 [nshared] smallest local dof index holding that node (0-based)
    neigh_rank: np.ndarray    # int32 [nneigh]  ascending
    neigh_off: np.ndarray     # int32 [nneigh+1]
    neigh_idx: np.ndarray     # int32 [neigh_off[-1]] indices into shared_*; ascending key per neighbour
    bnd_elem: np.ndarray      # int32 elements owning at least one shared node (0-based)

    @property
    def nshared(self):
        return int(self.shared_key.size)


def _unique_first(keys_flat, cand_flat):
    """Sorted unique candidate keys and, for each, the smallest local dof index that carries it."""
    idx = torch.nonzero(cand_flat, as_tuple=False).view(-1)
    k = keys_flat[idx]
    # stable sort by key keeps ascending dof order inside runs of equal keys
    ks, order = torch.sort(k, stable=True)
    idx = idx[order]
    if ks.numel() == 0:
        return ks, idx
    first = torch.ones_like(ks, dtype=torch.bool)
    first[1:] = ks[1:] != ks[:-1]
    return ks[first], idx[first]


def find_shared_nodes(keys, candidates, lxyz, rank=None, nranks=None, group=None):
    """keys: int64 tensor of n = nelv*lxyz global node ids; candidates: bool tensor of the same shape.
    Collective over `group` (every rank must call it).  Works on CPU (gloo) or CUDA (nccl) tensors."""
    if nranks is None:
        nranks = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if nranks > 1 else 0
    kf = keys.reshape(-1)
    cf = candidates.reshape(-1)
    uk, udof = _unique_first(kf, cf)
    empty = SharedNodes(np.zeros(0, np.int64), np.zeros(0, np.int32), np.zeros(0, np.int32),
                        np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32))
    if nranks == 1:
        return empty
    dev = kf.device
    # all-gather the (padded) unique candidate keys
    cnt = torch.tensor([uk.numel()], dtype=torch.int64, device=dev)
    cnts = [torch.zeros_like(cnt) for _ in range(nranks)]
    dist.all_gather(cnts, cnt, group=group)
    cnts = [int(c.item()) for c in cnts]
    m = max(max(cnts), 1)
    pad = torch.full((m,), -1, dtype=torch.int64, device=dev)
    pad[:uk.numel()] = uk
    allk = [torch.empty_like(pad) for _ in range(nranks)]
    dist.all_gather(allk, pad, group=group)

    with_rank = []                     # (rank r, bool mask over uk: shared with r)
    any_shared = torch.zeros(uk.numel(), dtype=torch.bool, device=dev)
    for r in range(nranks):
        if r == rank or cnts[r] == 0 or uk.numel() == 0:
            continue
        other = allk[r][:cnts[r]]      # sorted ascending
        pos = torch.searchsorted(other, uk).clamp_(max=cnts[r] - 1)
        hit = other[pos] == uk
        if bool(hit.any()):
            with_rank.append((r, hit))
            any_shared |= hit
    if not with_rank:
        return empty
    # compact numbering of the shared nodes (ascending key)
    sidx = torch.cumsum(any_shared.to(torch.int64), 0) - 1
    shared_key = uk[any_shared]
    shared_dof = udof[any_shared]
    neigh_rank, neigh_off, neigh_idx = [], [0], []
    for r, hit in with_rank:
        neigh_rank.append(r)
        neigh_idx.append(sidx[hit])
        neigh_off.append(neigh_off[-1] + int(hit.sum().item()))
    neigh_idx = torch.cat(neigh_idx)
    # elements owning a shared node: every local dof whose key is a shared key
    cand_idx = torch.nonzero(cf, as_tuple=False).view(-1)
    ck = kf[cand_idx]
    pos = torch.searchsorted(shared_key, ck).clamp_(max=shared_key.numel() - 1)
    is_sh = shared_key[pos] == ck
    bnd = torch.unique(torch.div(cand_idx[is_sh], lxyz, rounding_mode="floor"))
    return SharedNodes(shared_key.cpu().numpy().astype(np.int64), shared_dof.cpu().numpy().astype(np.int32),
                       np.asarray(neigh_rank, dtype=np.int32), np.asarray(neigh_off, dtype=np.int32),
                       neigh_idx.cpu().numpy().astype(np.int32), bnd.cpu().numpy().astype(np.int32))
