"""Regenerates tests/golden/debugging_pipe_mesh.npz from the reference's mesh fixture
/root/reference/data/debugging_pipe.nmsh (little-endian; int32 nelv, gdim; nelv records
{int32 el_idx; 8 x {int32 v_idx; float64 xyz[3]}}; then zones and curves, not needed here).

Only the hexahedra (vertex ids and coordinates, in the file's own element and vertex order) are kept: that
is what the adjoint-RHS path needs to build nodes, geometric factors and the gather-scatter map on a mesh
whose numbering is not ours.  Run in the build container (the GPU box has no /root/reference)."""
import os
import struct

import numpy as np

SRC = "/root/reference/data/debugging_pipe.nmsh"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "debugging_pipe_mesh.npz")


def read_nmsh_hexes(path):
    b = open(path, "rb").read()
    nelv, gdim = struct.unpack_from("<ii", b, 0)
    assert gdim == 3
    off = 8
    vid = np.zeros((nelv, 8), dtype=np.int32)
    xyz = np.zeros((nelv, 8, 3), dtype=np.float64)
    for e in range(nelv):
        off += 4                                   # el_idx
        for v in range(8):
            vid[e, v] = struct.unpack_from("<i", b, off)[0]
            xyz[e, v] = struct.unpack_from("<ddd", b, off + 4)
            off += 28
    return vid, xyz


if __name__ == "__main__":
    vid, xyz = read_nmsh_hexes(SRC)
    np.savez_compressed(DST, vertex_id=vid, vertex_xyz=xyz)
    print(f"{DST}: {vid.shape[0]} hexahedra, {np.unique(vid).size} vertices")
