"""ctypes loader for libneko_top_b200.so (the C ABI in include/neko_top_b200.h).

There is no CPU fallback: if the shared library is missing the import of any compute entry point
raises, and the library itself aborts/returns an error when no CUDA device is present."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libneko_top_b200.so")

# every symbol include/neko_top_b200.h declares (checked by tests/test_abi.py against the header)
SYMBOLS = [
    "b200_version", "b200_set_abort_on_error", "b200_last_error", "b200_launch_count",
    "b200_adjrhs_create", "b200_adjrhs_free", "b200_adjrhs_set_stream", "b200_adjrhs_set_space",
    "b200_adjrhs_set_geometry", "b200_adjrhs_set_params", "b200_adjrhs_set_lube_mask",
    "b200_adjrhs_compute", "b200_adjrhs_step", "b200_adjrhs_step_host",
    "b200_adv_adjoint_compute", "b200_adv_linear_compute", "b200_adv_dealias_init",
    "b200_adv_adjoint_dealias_compute", "b200_adv_linear_dealias_compute", "b200_adjrhs_set_dealias",
    "b200_brinkman_compute",
    "b200_lube_compute", "b200_opcolv", "b200_ramp_forward", "b200_ramp_backward",
    "b200_sensitivity", "b200_steady_field_update",
    "b200_curl", "b200_curlcurl_forcing", "b200_min_dissipation_objective", "b200_mask_exterior_const", "b200_pde_filter_apply",
    "b200_sumab", "b200_makeabf", "b200_makebdf", "b200_makeabf_bdf",
    "b200_gs_init", "b200_gs_get_classes", "b200_gs_op", "b200_gs_op3",
    "b200_comm_unique_id", "b200_comm_init", "b200_gs_init_shared", "b200_gs_init_shared_from_keys",
    "b200_adjrhs_set_boundary_elements", "b200_adjrhs_set_element_order", "b200_adjrhs_gs_info", "b200_adjrhs_set_gs_fused", "b200_adjrhs_set_xstage", "b200_adjrhs_xstage_info", "b200_adjrhs_enable_timing", "b200_adjrhs_get_timing", "b200_adjrhs_get_phase_timing",
]


def build(force=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    csrc = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc)]
    srcs.append(os.path.join(_HERE, "..", "include", "neko_top_b200.h"))
    if (not force and os.path.exists(SO_PATH)
            and all(os.path.getmtime(SO_PATH) >= os.path.getmtime(s) for s in srcs)):
        return SO_PATH
    subprocess.check_call(["make", "-j2", "-C", csrc], stdout=subprocess.DEVNULL)
    return SO_PATH


_lib = None


def lib():
    """Load the library (raises if it has not been built: no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C neko-top_b200/csrc).  neko_top_b200 has no CPU fallback.")
        L = C.CDLL(SO_PATH, mode=C.RTLD_GLOBAL)
        L.b200_last_error.restype = C.c_char_p
        L.b200_launch_count.restype = C.c_int64
        for s in SYMBOLS:
            getattr(L, s)   # AttributeError if the ABI and this table disagree
        _lib = L
    return _lib


class B200Error(RuntimeError):
    pass


def check(status):
    if status != 0:
        raise B200Error(lib().b200_last_error().decode())


def set_abort_on_error(flag):
    lib().b200_set_abort_on_error(C.byref(C.c_int(int(flag))))
