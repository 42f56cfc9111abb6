"""ctypes binding of libsimples_b200.so (include/simples_b200.h).

The CUDA library is the product.  There is no CPU fallback: if the shared
library is missing this module raises, and if no sm_100 device is usable every
compute entry point returns SIMPLES_ENODEV which is raised as SimplesError.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsimples_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)

ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)

COMPAT_STALE_DT = 1
COMPAT_PI = 2
COMPAT_MU_DIV_N = 4
COMPAT_PRIORB_PREV = 8
COMPAT_DEFAULT = 15


class SimplesError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"simples_b200 error {code}: {msg}")
        self.code = code


class EmArgs(C.Structure):
    _fields_ = [
        ("G", C.c_int32), ("n", C.c_int32), ("M0", C.c_int32), ("K0", C.c_int32), ("MB", C.c_int32),
        ("Y", C.c_void_p), ("Y0", C.c_void_p), ("pg", C.c_void_p),
        ("cutoff", C.c_double), ("iter", C.c_int32),
        ("beta", C.c_void_p), ("sigma", C.c_void_p), ("lambda_", C.c_void_p), ("pi", C.c_void_p),
        ("z", C.c_void_p), ("mu", C.c_void_p), ("celltype", C.c_void_p),
        ("penl", C.c_double),
        ("est_z", C.c_int32), ("max_lambda", C.c_int32), ("est_lam", C.c_int32), ("impt_it", C.c_int32),
        ("sigma0", C.c_double), ("pi_alpha", C.c_double),
        ("num_mc", C.c_int32),
        ("lower", C.c_double), ("upper", C.c_double),
        ("seed", C.c_uint64), ("ref_compat_flags", C.c_uint32),
        ("n_total", C.c_int64), ("cell_offset", C.c_int64),
        ("allreduce", ALLREDUCE_FN), ("allreduce_user", C.c_void_p), ("stream", C.c_void_p),
    ]


class EmOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in
                ("loglik", "pi", "mu", "sigma", "beta", "lambda_", "z", "Ef", "Varf", "Y", "geneM", "geneSd", "priorB")]


class MiArgs(C.Structure):
    _fields_ = [
        ("G", C.c_int32), ("n", C.c_int32), ("M0", C.c_int32), ("K0", C.c_int32), ("MB", C.c_int32),
        ("dat", C.c_void_p), ("Y", C.c_void_p),
        ("beta", C.c_void_p), ("lambda_", C.c_void_p), ("sigma", C.c_void_p), ("mu", C.c_void_p), ("pi", C.c_void_p),
        ("pos_mean", C.c_void_p), ("pos_sd", C.c_void_p), ("celltype", C.c_void_p),
        ("mcmc", C.c_int32), ("burnin", C.c_int32),
        ("pg", C.c_void_p), ("cutoff", C.c_double),
        ("seed", C.c_uint64), ("ref_compat_flags", C.c_uint32), ("want_consensus", C.c_int32),
        ("n_total", C.c_int64), ("cell_offset", C.c_int64),
        ("allreduce", ALLREDUCE_FN), ("allreduce_user", C.c_void_p), ("stream", C.c_void_p),
    ]


class MiOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("loglik", "impt", "impt_var", "EF", "varF", "consensus_cluster")]


class LqArgs(C.Structure):
    _fields_ = [
        ("Gq", C.c_int32), ("n", C.c_int32), ("M0", C.c_int32), ("K0", C.c_int32), ("MB", C.c_int32),
        ("dat", C.c_void_p), ("resp", C.c_void_p), ("sigma", C.c_void_p), ("pg", C.c_void_p),
        ("z", C.c_void_p), ("Ef", C.c_void_p), ("Varf", C.c_void_p), ("celltype", C.c_void_p),
        ("penl", C.c_double), ("cutoff", C.c_double),
        ("resid_mode", C.c_int32), ("impute", C.c_int32),
        ("seed", C.c_uint64), ("sweep", C.c_uint32),
        ("n_total", C.c_int64), ("cell_offset", C.c_int64),
        ("allreduce", ALLREDUCE_FN), ("allreduce_user", C.c_void_p), ("stream", C.c_void_p),
    ]


class LqOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("mu", "beta", "sigma", "Y")]


class InitArgs(C.Structure):
    _fields_ = [
        ("G", C.c_int32), ("n", C.c_int32), ("M0", C.c_int32),
        ("Y2", C.c_void_p), ("clus", C.c_void_p), ("pg1", C.c_void_p),
        ("p_min", C.c_double), ("cutoff", C.c_double),
        ("seed", C.c_uint64), ("sweep", C.c_uint32),
        ("n_total", C.c_int64), ("cell_offset", C.c_int64),
        ("allreduce", ALLREDUCE_FN), ("allreduce_user", C.c_void_p), ("stream", C.c_void_p),
    ]


class InitOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("impute", "pg", "mu", "sd")]


class ClusArgs(C.Structure):
    _fields_ = [
        ("G", C.c_int32), ("n", C.c_int32), ("K", C.c_int32), ("M0", C.c_int32),
        ("X", C.c_void_p), ("nstart", C.c_int32), ("iter_max", C.c_int32),
        ("seed", C.c_uint64), ("n_total", C.c_int64), ("cell_offset", C.c_int64),
        ("allreduce", ALLREDUCE_FN), ("allreduce_user", C.c_void_p), ("stream", C.c_void_p),
    ]


class ClusOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("clus", "cellmat", "d", "tot_withinss")]


EXPORTS = [
    "simples_abi_version", "simples_init", "simples_shutdown", "simples_last_error",
    "simples_em_impute", "simples_do_impute", "simples_lq_fit", "simples_init_impute", "simples_init_cluster", "simples_em_create", "simples_em_run", "simples_em_fetch",
    "simples_em_sync", "simples_em_iter_ms", "simples_em_set_profile", "simples_em_profile",
    "simples_kernel_name", "simples_kernel_count", "simples_em_nnz0", "simples_em_kernel_launches",
    "simples_em_yimp0", "simples_bic",
    "simples_ctx_destroy",
    "simples_device_count", "simples_comm_unique_id", "simples_comm_init_rank", "simples_comm_destroy", "simples_comm_nranks",
    "simples_nccl_version", "simples_pack_mask",
]

_lib = None


def _point_at_nccl():
    """The library dlopens libnccl on demand (multi-GPU only).  Prefer the copy bundled with PyTorch (the one torch itself
    loads, so both share it) over a system libnccl; SIMPLES_NCCL_LIB set by the user wins."""
    if os.environ.get("SIMPLES_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            cand = os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["SIMPLES_NCCL_LIB"] = cand
    except Exception:  # noqa: BLE001
        pass


def load():
    """Load the CUDA library; raise (never fall back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(simples_b200 has no CPU fallback)")
    _point_at_nccl()
    lib = C.CDLL(LIB_PATH)
    lib.simples_abi_version.restype = C.c_int
    lib.simples_comm_unique_id.argtypes = [C.c_void_p, C.c_size_t]
    lib.simples_comm_init_rank.argtypes = [C.c_int, C.c_int, C.c_void_p]
    lib.simples_comm_destroy.restype = None
    lib.simples_pack_mask.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_void_p]
    lib.simples_pack_mask.restype = C.c_int64
    lib.simples_init.argtypes = [C.c_int, C.POINTER(C.c_int)]
    lib.simples_last_error.restype = C.c_char_p
    lib.simples_em_impute.argtypes = [C.POINTER(EmArgs), C.POINTER(EmOut)]
    lib.simples_do_impute.argtypes = [C.POINTER(MiArgs), C.POINTER(MiOut)]
    lib.simples_lq_fit.argtypes = [C.POINTER(LqArgs), C.POINTER(LqOut)]
    lib.simples_init_impute.argtypes = [C.POINTER(InitArgs), C.POINTER(InitOut)]
    lib.simples_init_cluster.argtypes = [C.POINTER(ClusArgs), C.POINTER(ClusOut)]
    lib.simples_em_create.argtypes = [C.POINTER(EmArgs), C.POINTER(C.c_void_p)]
    lib.simples_em_run.argtypes = [C.c_void_p, C.c_int]
    lib.simples_em_fetch.argtypes = [C.c_void_p, C.POINTER(EmOut)]
    lib.simples_em_sync.argtypes = [C.c_void_p]
    lib.simples_em_iter_ms.argtypes = [C.c_void_p, c_double_p, C.c_int]
    lib.simples_em_set_profile.argtypes = [C.c_void_p, C.c_int]
    lib.simples_em_profile.argtypes = [C.c_void_p, C.POINTER(C.c_int64), c_double_p, C.c_int]
    lib.simples_kernel_name.argtypes = [C.c_int]
    lib.simples_kernel_name.restype = C.c_char_p
    lib.simples_kernel_count.restype = C.c_int
    lib.simples_em_nnz0.argtypes = [C.c_void_p]
    lib.simples_em_nnz0.restype = C.c_int64
    lib.simples_em_kernel_launches.argtypes = [C.c_void_p]
    lib.simples_em_kernel_launches.restype = C.c_int64
    lib.simples_em_yimp0.argtypes = [C.c_void_p, C.c_void_p]
    lib.simples_bic.argtypes = [C.c_double, C.c_int64, C.c_int32, C.c_int32, C.c_int32, c_double_p, c_double_p]
    lib.simples_ctx_destroy.argtypes = [C.c_void_p]
    lib.simples_ctx_destroy.restype = None
    _lib = lib
    return lib


def check(rc):
    if rc < 0:
        raise SimplesError(rc, load().simples_last_error().decode("utf-8", "replace"))
    return rc


def init(device=0):
    """``init(i)``: bind this process to GPU i (one process per GPU).  ``init([i, j, ...])``: this process drives all
    listed GPUs -- EM_impute / do_impute then shard the cells over them inside the library (host threads + NCCL)."""
    ids = [int(d) for d in device] if isinstance(device, (list, tuple)) else [int(device)]
    devs = (C.c_int * len(ids))(*ids)
    check(load().simples_init(len(ids), devs))


def comm_init_from_torch(group=None):
    """One process per GPU (torchrun): build the library's own NCCL rank communicator, the unique id travelling over
    torch.distributed.  Afterwards sharded calls need no Python all-reduce callback."""
    import torch.distributed as dist
    lib = load()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    buf = C.create_string_buffer(128)
    if rank == 0:
        check(lib.simples_comm_unique_id(buf, 128))
    obj = [bytes(buf.raw)]
    dist.broadcast_object_list(obj, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    idb = C.create_string_buffer(obj[0], 128)
    check(lib.simples_comm_init_rank(world, rank, idb))
    return world
