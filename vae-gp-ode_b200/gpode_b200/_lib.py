"""ctypes binding of libgpode.so (C ABI declared in include/gpode.h).  No torch types cross the ABI:
only device pointers, sizes and the CUDA stream handle."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libgpode.so")

RBF_SHARED, RBF_DIMWISE, DF = 0, 1, 2
EULER, MIDPOINT, RK4 = 0, 1, 2
METHODS = {"euler": EULER, "midpoint": MIDPOINT, "rk4": RK4}
STAGES = {EULER: 1, MIDPOINT: 2, RK4: 4}
VARIANTS = {"rbf_shared": RBF_SHARED, "rbf_dimwise": RBF_DIMWISE, "df": DF}

_fp = ctypes.c_void_p


class GpodeProblem(ctypes.Structure):
    _fields_ = [("variant", ctypes.c_int32), ("L", ctypes.c_int32), ("N", ctypes.c_int32), ("D_in", ctypes.c_int32),
                ("D_out", ctypes.c_int32), ("M", ctypes.c_int32), ("S", ctypes.c_int32), ("flags", ctypes.c_int32),
                ("Z", _fp), ("ell", _fp), ("var", _fp), ("eps", _fp), ("phase", _fp), ("w", _fp), ("nu", _fp), ("B", _fp)]


class GpodeParamGrads(ctypes.Structure):
    _fields_ = [("d_Z", _fp), ("d_ell", _fp), ("d_var", _fp), ("d_nu", _fp), ("d_B", _fp)]


EXPORTS = ["gpode_version", "gpode_error_string", "gpode_forward_kernel", "gpode_cluster_size", "gpode_workspace_bytes", "gpode_rollout_save_floats",
           "gpode_field_fwd", "gpode_field_bwd", "gpode_rollout_fwd", "gpode_rollout_bwd",
           "gpode_nu_workspace_bytes", "gpode_nu_save_floats", "gpode_compute_nu_fwd", "gpode_compute_nu_bwd",
           "gpode_inducing_sample_fwd", "gpode_inducing_sample_bwd", "gpode_kl_fwd", "gpode_kl_bwd",
           "gpode_philox_fill", "gpode_philox_raw", "gpode_bernoulli_workspace_bytes", "gpode_bernoulli_lhood_fwd", "gpode_bernoulli_lhood_bwd"]

_lib = None


def load():
    """Load libgpode.so; fails loudly when it has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libgpode.so not found at %s -- build it with `make -C vae-gp-ode_b200/csrc -j8` "
                           "(or python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    P, PG = ctypes.POINTER(GpodeProblem), ctypes.POINTER(GpodeParamGrads)
    i32, sz = ctypes.c_int, ctypes.c_size_t
    lib.gpode_version.restype = i32
    lib.gpode_version.argtypes = []
    lib.gpode_error_string.restype = ctypes.c_char_p
    lib.gpode_error_string.argtypes = [i32]
    lib.gpode_forward_kernel.restype = i32
    lib.gpode_forward_kernel.argtypes = [P]
    lib.gpode_cluster_size.restype = i32
    lib.gpode_cluster_size.argtypes = [P]
    lib.gpode_workspace_bytes.restype = sz
    lib.gpode_workspace_bytes.argtypes = [P, i32, i32]
    lib.gpode_rollout_save_floats.restype = sz
    lib.gpode_rollout_save_floats.argtypes = [P, i32, i32]
    lib.gpode_field_fwd.restype = i32
    lib.gpode_field_fwd.argtypes = [P, _fp, _fp, _fp, _fp, sz, _fp]
    lib.gpode_field_bwd.restype = i32
    lib.gpode_field_bwd.argtypes = [P, _fp, _fp, _fp, _fp, _fp, PG, _fp, sz, _fp]
    lib.gpode_rollout_fwd.restype = i32
    lib.gpode_rollout_fwd.argtypes = [P, _fp, i32, _fp, i32, i32, i32, _fp, _fp, _fp, sz, _fp]
    lib.gpode_rollout_bwd.restype = i32
    lib.gpode_rollout_bwd.argtypes = [P, _fp, i32, i32, i32, _fp, _fp, _fp, _fp, PG, _fp, sz, _fp]
    lib.gpode_nu_workspace_bytes.restype = sz
    lib.gpode_nu_workspace_bytes.argtypes = [P]
    lib.gpode_nu_save_floats.restype = sz
    lib.gpode_nu_save_floats.argtypes = [P]
    lib.gpode_compute_nu_fwd.restype = i32
    lib.gpode_compute_nu_fwd.argtypes = [P, _fp, _fp, _fp, _fp, _fp, _fp, sz, _fp]
    lib.gpode_compute_nu_bwd.restype = i32
    lib.gpode_compute_nu_bwd.argtypes = [P, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, sz, _fp]
    lib.gpode_inducing_sample_fwd.restype = i32
    lib.gpode_inducing_sample_fwd.argtypes = [i32, i32, i32, _fp, _fp, _fp, _fp, _fp]
    lib.gpode_inducing_sample_bwd.restype = i32
    lib.gpode_inducing_sample_bwd.argtypes = [i32, i32, i32, _fp, _fp, _fp, _fp, _fp]
    lib.gpode_kl_fwd.restype = i32
    lib.gpode_kl_fwd.argtypes = [i32, i32, _fp, _fp, _fp, _fp]
    lib.gpode_kl_bwd.restype = i32
    lib.gpode_kl_bwd.argtypes = [i32, i32, _fp, _fp, _fp, _fp, _fp, _fp]
    u64 = ctypes.c_uint64
    lib.gpode_philox_fill.restype = i32
    lib.gpode_philox_fill.argtypes = [i32, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(u64), ctypes.POINTER(ctypes.c_int32), u64, u64, _fp]
    lib.gpode_philox_raw.restype = i32
    lib.gpode_philox_raw.argtypes = [_fp, _fp, _fp, i32, _fp]
    lib.gpode_bernoulli_workspace_bytes.restype = sz
    lib.gpode_bernoulli_workspace_bytes.argtypes = [i32]
    lib.gpode_bernoulli_lhood_fwd.restype = i32
    lib.gpode_bernoulli_lhood_fwd.argtypes = [i32, i32, ctypes.c_int64, _fp, _fp, _fp, _fp, sz, _fp]
    lib.gpode_bernoulli_lhood_bwd.restype = i32
    lib.gpode_bernoulli_lhood_bwd.argtypes = [i32, i32, ctypes.c_int64, _fp, _fp, _fp, _fp, _fp]
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().gpode_error_string(int(rc)).decode()
        raise RuntimeError("%s failed: %s (code %d)" % (what, msg, rc))


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_handle(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("gpode_b200 runs on CUDA devices only (got a %s tensor); there is no CPU path" % t.device)
        if t.dtype != torch.float32:
            raise RuntimeError("gpode_b200 computes in fp32 (got %s)" % t.dtype)


# GpodeProblem.flags (include/gpode.h): kernel selection for the parity tests / A-B measurements; 0 = shape-driven defaults
FLAG_FWD_MMA, FLAG_FWD_TCGEN05, FLAG_BWD_MMA, FLAG_BWD_TCGEN05 = 1, 2, 4, 16
_flags = 0


class kernel_flags:
    """``with kernel_flags(FLAG_FWD_MMA): ...`` -- every problem built inside the block carries these flags (an explicit field
    of the C ABI; the library itself reads no environment variable and keeps no state)."""

    def __init__(self, flags):
        self.flags = int(flags)

    def __enter__(self):
        global _flags
        self.prev, _flags = _flags, self.flags
        return self

    def __exit__(self, *exc):
        global _flags
        _flags = self.prev
        return False


def make_problem(variant, L, N, D_in, D_out, M, S, Z, ell, var, eps, phase, w, nu, B=None, flags=None):
    p = GpodeProblem()
    p.variant, p.L, p.N, p.D_in, p.D_out, p.M, p.S = variant, L, N, D_in, D_out, M, S
    p.flags = _flags if flags is None else int(flags)
    p.Z, p.ell, p.var, p.eps, p.phase, p.w, p.nu = ptr(Z), ptr(ell), ptr(var), ptr(eps), ptr(phase), ptr(w), ptr(nu)
    p.B = ptr(B)
    return p


def workspace(p, T, method, device):
    nbytes = load().gpode_workspace_bytes(ctypes.byref(p), T, method)
    if nbytes == 0:
        raise RuntimeError("gpode: unsupported problem (variant=%d L=%d N=%d D_in=%d D_out=%d M=%d S=%d): "
                           "D_in/D_out <= 16, DF D <= 8, M <= 512" % (p.variant, p.L, p.N, p.D_in, p.D_out, p.M, p.S))
    return torch.empty(nbytes, dtype=torch.uint8, device=device), nbytes
