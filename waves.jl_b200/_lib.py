"""ctypes binding of libwaves_b200.so -- the C ABI declared in include/waves_b200.h.

There is no CPU fallback: if the CUDA library is missing or cannot be loaded this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# WAVES_B200_LIB selects another build of the same ABI (scripts/tune_build.sh); there is still no CPU fallback
SO_PATH = os.environ.get("WAVES_B200_LIB") or os.path.join(_HERE, "libwaves_b200.so")
CSRC = os.path.join(_HERE, "csrc")

fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int32)


class WavesConfig(C.Structure):
    """struct waves_config (include/waves_b200.h)."""
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("n_env", C.c_int32), ("device", C.c_int32),
        ("c0", C.c_float), ("dt", C.c_float), ("pml_width", C.c_float), ("pml_scale", C.c_float),
        ("x", fp), ("y", fp), ("sigma", fp), ("grad8", fp),
        ("d_omega", C.c_float), ("ny_global", C.c_int32), ("row0", C.c_int32), ("flags", C.c_uint32),
    ]


class PeerInfo(C.Structure):
    """struct waves_peer_info."""
    _fields_ = [("ipc_u0", C.c_ubyte * 64), ("ipc_u1", C.c_ubyte * 64), ("ipc_flags", C.c_ubyte * 64),
                ("ny_alloc", C.c_int32), ("ny_own0", C.c_int32), ("ny_own", C.c_int32), ("device", C.c_int32),
                ("reserved", C.c_int32 * 12)]


class HaloDesc(C.Structure):
    """struct waves_halo_desc."""
    _fields_ = [
        ("send_lo", C.c_void_p), ("send_hi", C.c_void_p), ("recv_lo", C.c_void_p), ("recv_hi", C.c_void_p),
        ("plane_stride", C.c_int64), ("n_planes", C.c_int32), ("block_floats", C.c_int32), ("pitch", C.c_int32),
    ]


class LatentConfig(C.Structure):
    """struct waves_latent_config."""
    _fields_ = [("n", C.c_int32), ("device", C.c_int32), ("c0", C.c_float), ("dt", C.c_float), ("pml_width", C.c_float),
                ("pml_scale", C.c_float), ("pml0", C.c_float), ("dx", C.c_float), ("x", fp), ("grad8", fp)]


# every symbol include/waves_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "waves_version": (C.c_int, []),
    "waves_build_flags": (C.c_int, []),
    "waves_last_error": (C.c_char_p, []),
    "waves_create": (C.c_int, [C.POINTER(WavesConfig), C.POINTER(C.c_void_p)]),
    "waves_destroy": (C.c_int, [C.c_void_p]),
    "waves_sync": (C.c_int, [C.c_void_p]),
    "waves_set_state": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "waves_get_state": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "waves_set_source": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_float]),
    "waves_set_design": (C.c_int, [C.c_void_p, C.c_int, C.c_int, fp, fp, C.c_float, C.c_float]),
    "waves_set_design_batch": (C.c_int, [C.c_void_p, C.c_int, fp, fp, C.c_float, C.c_float]),
    "waves_set_speed_field": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "waves_rhs": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_void_p]),
    "waves_step": (C.c_int, [C.c_void_p, C.c_float, C.c_int]),
    "waves_integrate": (C.c_int, [C.c_void_p, fp, C.c_int, C.c_int, C.c_void_p, ip, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "waves_set_traj_stride": (C.c_int, [C.c_void_p, C.c_int]),
    "waves_set_graph": (C.c_int, [C.c_void_p, C.c_int]),
    "waves_observe": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "waves_energy": (C.c_int, [C.c_void_p, C.c_void_p]),
    "waves_adjoint": (C.c_int, [C.c_void_p, fp, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p]),
    "waves_set_adjoint_checkpoint": (C.c_int, [C.c_void_p, C.c_int]),
    "waves_halo_describe": (C.c_int, [C.c_void_p, C.POINTER(HaloDesc)]),
    "waves_halo_pack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "waves_halo_unpack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "waves_peer_export": (C.c_int, [C.c_void_p, C.POINTER(PeerInfo)]),
    "waves_peer_attach": (C.c_int, [C.c_void_p, C.POINTER(PeerInfo), C.POINTER(PeerInfo)]),
    "waves_stream": (C.c_void_p, [C.c_void_p]),
    "waves_range_f32": (C.c_int, [C.c_float, C.c_float, C.c_int, fp]),
    "waves_build_pml_profile": (C.c_int, [fp, C.c_int, C.c_float, C.c_float, fp]),
    "waves_build_gradient8": (C.c_int, [fp, C.c_int, fp]),
    "waves_build_normal": (C.c_int, [fp, C.c_int, fp, C.c_int, C.c_int, fp, fp, fp, fp]),
    "waves_mean_diff": (C.c_float, [fp, C.c_int]),
    "waves_latent_create": (C.c_int, [C.POINTER(LatentConfig), C.POINTER(C.c_void_p)]),
    "waves_latent_destroy": (C.c_int, [C.c_void_p]),
    "waves_latent_integrate": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "waves_latent_adjoint": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    "waves_latent_build_pml": (C.c_int, [fp, C.c_int, C.c_float, C.c_float, fp, fp]),
    "waves_latent_launch_count": (C.c_int64, [C.c_void_p]),
    "waves_latent_last_kernel_ms": (C.c_float, [C.c_void_p]),
    "waves_latent_set_generic": (C.c_int, [C.c_void_p, C.c_int]),
    "waves_latent_set_variant": (C.c_int, [C.c_void_p, C.c_int]),
    "waves_launch_count": (C.c_int64, [C.c_void_p]),
    "waves_graph_replays": (C.c_int64, [C.c_void_p]),
    "waves_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "waves_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
}

_LIB = None


def build(force: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libwaves_b200.so (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    srcs.append(os.path.join(_HERE, "..", "include", "waves_b200.h"))
    stale = (not os.path.exists(SO_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(SO_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", CSRC] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return SO_PATH


def lib():
    """The loaded library.  Raises (never falls back) if it is absent."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(waves.jl_b200 has no CPU fallback)")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


class WavesError(RuntimeError):
    pass


def check(rc: int):
    if rc != 0:
        raise WavesError(lib().waves_last_error().decode("utf-8", "replace"))
