"""ctypes binding of librimu_b200.so (include/rimu_b200.h) and its nvcc build recipe.

There is no CPU fallback: if the shared library is missing or no CUDA device is present,
every compute entry point raises `RimuB200Error`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("RIMU_B200_LIB") or os.path.join(_HERE, "librimu_b200.so")  # override: kernel-tuning builds
CSRC = os.path.join(_HERE, "csrc")
SOURCES = ["api.cu", "sort.cu", "sector.cu", "step_hk.cu"]   # step_hk.cu is compiled once per HamKind (-DRIMU_HK=n), in parallel
HEADERS = ["common.cuh", "hamiltonians.cuh", "kernels.cuh", "partition.cuh", "ham_host.h", "step_math.cuh", "internal.cuh", "sector.cuh"]
NUM_HAM_KINDS = 10
OBJ_DIR = os.environ.get("RIMU_B200_OBJ_DIR") or os.path.join("/tmp", "rimu_b200_build_" + str(os.getuid()))  # objects stay out of the tree

NVCC_FLAGS = [
    "-std=c++17", "-O3", "--fmad=false", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC",
]

MAX_MODES = 128
MAX_TABLE_MODES = 64

OK, ERR_TABLE_FULL, ERR_VECTOR_FULL, ERR_EXCHANGE_FULL, ERR_WORKMEM = 0, 1, 2, 3, 4
ERR_INVALID, ERR_CUDA, ERR_NCCL, ERR_NO_DEVICE = -1, -2, -3, -4
ADDR_BOSE, ADDR_FERMI, ADDR_FERMI2C, ADDR_COMPOSITE = 0, 1, 2, 3
MAX_COMPONENTS = 4
HUBBARD_REAL_1D, HUBBARD_MOM_1D, HUBBARD_REAL_SPACE, TRANSCORRELATED_1D = 0, 1, 2, 3
HUBBARD_REAL_1D_EP, EXTENDED_HUBBARD_REAL_1D = 4, 5
EXTENDED_HUBBARD_MOM_1D, HUBBARD_MOM_1D_EP = 6, 7
BC_PERIODIC, BC_HARD_WALL, BC_TWISTED = 0, 1, 2
VAL_F64, VAL_I64 = 0, 1
STYLE_DETERMINISTIC, STYLE_INTEGER, STYLE_SEMISTOCHASTIC, STYLE_WITH_THRESHOLD = 0, 1, 2, 3
ANNIHILATE_HASH, ANNIHILATE_SORT, ANNIHILATE_PARTITION = 0, 1, 2


class RimuB200Error(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"[rimu_b200 status {status}] {message}")
        self.status = status


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(_ROOT, "include", "rimu_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None, kinds=None) -> str:
    """Compile every CUDA source for sm_100a into rimu.jl_b200/librimu_b200.so (in-tree).

    api.cu, sort.cu and one object per Hamiltonian kind (step_hk.cu with -DRIMU_HK=n) are compiled in parallel and linked
    with nvcc -shared.  `defines`/`out`/`kinds` are for kernel-tuning builds (scratch/)."""
    target = out or LIB_PATH
    if out is None and (os.environ.get("RIMU_B200_LIB") or (not force and not _stale())):
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    tag = "" if out is None else "_" + os.path.splitext(os.path.basename(out))[0]
    os.makedirs(OBJ_DIR, exist_ok=True)
    base = [nvcc] + NVCC_FLAGS + list(defines) + (["-Xptxas", "-v"] if verbose else [])
    jobs = []
    for s in ("api.cu", "sort.cu", "sector.cu"):
        jobs.append((base + ["-c", os.path.join(CSRC, s), "-o", os.path.join(OBJ_DIR, s[:-3] + tag + ".o")]))
    for hk in (range(NUM_HAM_KINDS) if kinds is None else kinds):
        jobs.append((base + [f"-DRIMU_HK={hk}", "-c", os.path.join(CSRC, "step_hk.cu"), "-o", os.path.join(OBJ_DIR, f"step_hk{hk}{tag}.o")]))

    def run(cmd):
        return cmd, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        results = list(ex.map(run, jobs))
    log = ""
    for cmd, res in results:
        if res.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        log += res.stderr
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + [c[c.index("-o") + 1] for c, _ in results] + ["-o", target, "-ldl"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(log)
    return target


class HamDesc(C.Structure):
    _fields_ = [
        ("model", C.c_int32), ("addr_kind", C.c_int32), ("num_modes", C.c_int32), ("num_components", C.c_int32),
        ("num_particles", C.c_int32 * 2),
        ("ndim", C.c_int32), ("dims", C.c_int32 * 3), ("fold", C.c_int32 * 3),
        ("cutoff", C.c_int32), ("three_body_term", C.c_int32), ("has_potential", C.c_int32), ("boundary_condition", C.c_int32),
        ("u", C.c_double), ("t", C.c_double), ("v", C.c_double),
        ("t_comp", C.c_double * 2), ("u_mat", C.c_double * 4),
        ("kes", C.c_double * MAX_TABLE_MODES), ("ws", C.c_double * MAX_TABLE_MODES), ("us", C.c_double * MAX_TABLE_MODES),
        ("potential", C.c_double * (2 * MAX_MODES)),
        ("comp_kind", C.c_int32 * MAX_COMPONENTS), ("comp_particles", C.c_int32 * MAX_COMPONENTS),
        ("comp_t", C.c_double * MAX_COMPONENTS), ("comp_u", C.c_double * (MAX_COMPONENTS * MAX_COMPONENTS)),
    ]


class StepParams(C.Structure):
    _fields_ = [
        ("style", C.c_int32), ("plain_h", C.c_int32),
        ("shift", C.c_double), ("time_step", C.c_double), ("boost", C.c_double),
        ("proj_threshold", C.c_double), ("rel_threshold", C.c_double), ("abs_threshold", C.c_double),
        ("compress_threshold", C.c_double),
        ("seed", C.c_uint64), ("step", C.c_uint64), ("table_slots", C.c_uint64),
        ("initiator_rule", C.c_int32), ("ordered", C.c_int32), ("initiator_threshold", C.c_double),
    ]


class StepStats(C.Structure):
    _fields_ = [
        ("exact_steps", C.c_int64), ("inexact_steps", C.c_int64), ("spawn_attempts", C.c_int64),
        ("len_before", C.c_int64), ("len", C.c_int64),
        ("spawns", C.c_double), ("deaths", C.c_double), ("clones", C.c_double), ("zombies", C.c_double),
        ("norm1", C.c_double),
        ("ispawns", C.c_int64), ("ideaths", C.c_int64), ("iclones", C.c_int64), ("izombies", C.c_int64),
        ("inorm1", C.c_int64),
        ("local_len", C.c_int64), ("sent_records", C.c_int64),
        ("deposits", C.c_int64),
        ("ms_diag", C.c_float), ("ms_spawn", C.c_float), ("ms_exchange", C.c_float), ("ms_compact", C.c_float),
        ("ms_total", C.c_float), ("ms_reduce", C.c_float),
        ("buckets", C.c_int64), ("max_bucket_fill", C.c_int64),
    ]

    def asdict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class ShiftParams(C.Structure):
    """rimu_shift_params: shift strategy + DefaultShiftParameters of rimu_advance"""
    _fields_ = [
        ("strategy", C.c_int32), ("shift_mode", C.c_int32),
        ("target_walkers", C.c_double), ("zeta", C.c_double), ("xi", C.c_double),
        ("shift", C.c_double), ("pnorm", C.c_double),
        ("max_length", C.c_int64),
    ]


class Projector(C.Structure):
    """rimu_projector: a frozen projector (host pairs) whose dot with the vector rimu_advance reports after every step"""
    _fields_ = [("keys", C.POINTER(C.c_uint64)), ("values", C.POINTER(C.c_double)), ("n", C.c_int64)]


MAX_PROJECTORS = 8
SHIFT_DONT_UPDATE, SHIFT_LOG_UPDATE, SHIFT_LOG_UPDATE_AFTER_TARGET, SHIFT_DOUBLE_LOG_UPDATE, SHIFT_DOUBLE_LOG_UPDATE_AFTER_TARGET = range(5)

# every symbol include/rimu_b200.h declares: name -> (restype, argtypes)
_vp, _u64p, _i64p, _f64p, _u32p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_uint32)
SYMBOLS = {
    "rimu_last_error": (C.c_char_p, []),
    "rimu_version": (C.c_int, []),
    "rimu_sizeof_ham_desc": (C.c_int, []),
    "rimu_sizeof_step_params": (C.c_int, []),
    "rimu_sizeof_step_stats": (C.c_int, []),
    "rimu_ctx_create": (C.c_int, [C.c_int, C.c_int, C.c_uint64, C.POINTER(_vp)]),
    "rimu_ctx_destroy": (C.c_int, [_vp]),
    "rimu_ctx_synchronize": (C.c_int, [_vp]),
    "rimu_ctx_make_current": (C.c_int, [_vp]),
    "rimu_ctx_table_slots": (C.c_int, [_vp, _u64p]),
    "rimu_ctx_resize_table": (C.c_int, [_vp, C.c_uint64]),
    "rimu_ctx_stream": (C.c_int, [_vp, C.POINTER(_vp)]),
    "rimu_ctx_launch_count": (C.c_int, [_vp, _u64p]),
    "rimu_ctx_set_method": (C.c_int, [_vp, C.c_int]),
    "rimu_ctx_get_method": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "rimu_host_alloc": (C.c_int, [C.c_uint64, C.POINTER(_vp)]),
    "rimu_host_free": (C.c_int, [_vp]),
    "rimu_comm_unique_id": (C.c_int, [_vp]),
    "rimu_comm_init": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_uint64]),
    "rimu_comm_rank": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "rimu_comm_capacity": (C.c_int, [_vp, _u64p, _u64p]),
    "rimu_comm_reserve": (C.c_int, [_vp, C.c_uint64]),
    "rimu_comm_p2p": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "rimu_comm_detach": (C.c_int, [_vp]),
    "rimu_sector_create": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "rimu_sector_destroy": (C.c_int, [_vp]),
    "rimu_sector_dim": (C.c_int, [_vp, _u64p]),
    "rimu_sector_rank": (C.c_int, [_vp, _u64p, C.c_int64, _i64p]),
    "rimu_sector_keys": (C.c_int, [_vp, C.c_int64, C.c_int64, _u64p]),
    "rimu_sector_vec_create": (C.c_int, [_vp, C.POINTER(_vp)]),
    "rimu_sector_vec_destroy": (C.c_int, [_vp, _vp]),
    "rimu_sector_vec_set": (C.c_int, [_vp, _vp, _i64p, _f64p, C.c_int64]),
    "rimu_sector_vec_get": (C.c_int, [_vp, _vp, C.c_int64, C.c_int64, _f64p]),
    "rimu_sector_vec_gather": (C.c_int, [_vp, _vp, _i64p, C.c_int64, _f64p]),
    "rimu_sector_mul": (C.c_int, [_vp, _vp, _vp, C.POINTER(C.c_float)]),
    "rimu_sector_axpby": (C.c_int, [_vp, C.c_double, _vp, C.c_double, _vp]),
    "rimu_sector_dot": (C.c_int, [_vp, _vp, _vp, _f64p]),
    "rimu_sector_from_vec": (C.c_int, [_vp, _vp, _vp]),
    "rimu_sector_to_vec": (C.c_int, [_vp, _vp, _vp]),
    "rimu_comm_allreduce_f64": (C.c_int, [_vp, _f64p, C.c_int]),
    "rimu_addr_owner": (C.c_int, [_u64p, C.c_int, C.c_int]),
    "rimu_addr_hash": (C.c_uint64, [_u64p, C.c_int]),
    "rimu_ham_create": (C.c_int, [C.POINTER(HamDesc), C.POINTER(_vp)]),
    "rimu_ham_destroy": (C.c_int, [_vp]),
    "rimu_ham_words": (C.c_int, [_vp]),
    "rimu_ham_diagonal": (C.c_int, [_vp, _vp, _u64p, C.c_int64, _f64p]),
    "rimu_ham_num_offdiagonals": (C.c_int, [_vp, _vp, _u64p, C.c_int64, _i64p]),
    "rimu_ham_offdiagonals": (C.c_int, [_vp, _vp, _u64p, C.c_int64, C.c_int64, _u64p, _f64p]),
    "rimu_vec_create": (C.c_int, [_vp, C.c_int, C.c_uint64, C.POINTER(_vp)]),
    "rimu_vec_destroy": (C.c_int, [_vp]),
    "rimu_vec_reserve": (C.c_int, [_vp, C.c_uint64]),
    "rimu_vec_clear": (C.c_int, [_vp]),
    "rimu_vec_length": (C.c_int, [_vp, _i64p]),
    "rimu_vec_capacity": (C.c_int, [_vp, _u64p]),
    "rimu_vec_buckets": (C.c_int, [_vp, C.POINTER(C.c_uint32)]),
    "rimu_vec_rebucket": (C.c_int, [_vp, C.c_uint32]),
    "rimu_vec_segments": (C.c_int, [_vp, _u64p, C.POINTER(C.c_uint32)]),
    "rimu_vec_upload": (C.c_int, [_vp, _u64p, _vp, C.c_int64]),
    "rimu_vec_assign": (C.c_int, [_vp, _u64p, _vp, C.c_int64]),
    "rimu_vec_download": (C.c_int, [_vp, _u64p, _vp, C.c_int64, _i64p]),
    "rimu_vec_copy": (C.c_int, [_vp, _vp]),
    "rimu_vec_get": (C.c_int, [_vp, _u64p, _vp]),
    "rimu_vec_norm": (C.c_int, [_vp, C.c_int, _f64p]),
    "rimu_vec_scale": (C.c_int, [_vp, C.c_double]),
    "rimu_vec_dot": (C.c_int, [_vp, _vp, _f64p]),
    "rimu_vec_dot_sparse": (C.c_int, [_vp, _u64p, _f64p, C.c_int64, _f64p]),
    "rimu_vec_axpby": (C.c_int, [C.c_double, _vp, C.c_double, _vp, _vp]),
    "rimu_annihilate": (C.c_int, [_vp, _u64p, _vp, C.c_int64, C.c_int]),
    "rimu_annihilate_device": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int, C.POINTER(C.c_float)]),
    "rimu_step": (C.c_int, [_vp, _vp, C.POINTER(StepParams), _vp, _vp, C.POINTER(StepStats)]),
    "rimu_advance": (C.c_int, [_vp, _vp, C.POINTER(StepParams), C.POINTER(ShiftParams), _vp, _vp, C.c_int64, C.POINTER(Projector), C.c_int32,
                               C.POINTER(StepStats), _f64p, _f64p, _i64p, C.POINTER(C.c_int32)]),
    "rimu_sizeof_shift_params": (C.c_int, []),
    "rimu_step_key": (None, [C.c_uint64, C.c_uint64, _u32p]),
    "rimu_philox4x32_10": (None, [_u32p, _u32p, _u32p]),
}

_lib = None


def lib():
    """Load the shared library (never builds implicitly on a box without nvcc sources changed)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RimuB200Error(ERR_NO_DEVICE, f"{LIB_PATH} is missing: run __graft_entry__.build() (nvcc) first; "
                                               "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int):
    if status != 0:
        msg = lib().rimu_last_error()
        raise RimuB200Error(status, msg.decode() if msg else "unknown error")
