"""neko-top_b200: B200-native (sm_100a) adjoint right-hand-side path for Neko-TOP.

Only what the hot path needs:
  csrc/          hand-written CUDA kernels + the C ABI (include/neko_top_b200.h)
  _lib.py        ctypes loader of libneko_top_b200.so -- fails loudly when the library is missing
  operators.py   host-side mirror of the reference plug-in interface (advection_adjoint_t,
                 simple_brinkman_source_term_t, RAMP_mapping_t, gs_t%op, steady_simcomp_t)
  sem.py         what Neko's space_t / coef_t provide (GLL/GL points, D, geometric factors)
  workloads.py   synthetic meshes and fields of BASELINE.json's configs
  fortran/       the iso_c_binding shim a Neko-TOP maintainer drops in
"""
__version__ = "0.1.0"
