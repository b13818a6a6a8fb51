"""ctypes binding of libtcrisk.so (include/tcrisk.h).  There is no CPU fallback: if the
library is missing or a call fails, the error is raised."""
import ctypes as C
import os

from .params import TcrParams, TcrYearStats

_HERE = os.path.dirname(os.path.abspath(__file__))
# TCR_LIB_PATH: an instrumented build of the same sources (scripts/probes), never a different implementation
LIB_PATH = os.environ.get("TCR_LIB_PATH") or os.path.join(_HERE, "libtcrisk.so")

SYMBOLS = (
    "tcr_create", "tcr_destroy", "tcr_last_error", "tcr_set_stream", "tcr_synchronize", "tcr_version",
    "tcr_upload_static", "tcr_upload_masks", "tcr_alloc_tables", "tcr_upload_month", "tcr_upload_months", "tcr_upload_month_dev",
    "tcr_env_interp", "tcr_integrate", "tcr_run_years", "tcr_seed_attempts", "tcr_set_tuning",
    "tcr_launch_count", "tcr_set_interp_variant", "tcr_host_alloc", "tcr_host_free",
    "tcr_set_timing", "tcr_kernel_time", "tcr_poi_vmax", "tcr_exceedance", "tcr_prepare_month",
    "tcr_wind_stats", "tcr_set_entropy_table", "tcr_thermo_month", "tcr_set_entropy_table_reversible", "tcr_thermo_month_reversible", "tcr_set_workspace_budget", "tcr_fourier_ring_nodes", "tcr_rhs_eval", "tcr_set_shard",
)

_lib = None

c_i32p = C.POINTER(C.c_int32)
c_u32p = C.POINTER(C.c_uint32)
c_f64p = C.POINTER(C.c_double)
c_f32p = C.POINTER(C.c_float)


# tcr_allreduce_fn (include/tcrisk.h): int fn(void* user, void* d_buf, int64 count, int dtype, int op, void* stream)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p)


class TcrError(RuntimeError):
    pass


def load():
    """Load libtcrisk.so (building it first if nvcc is available and it is missing/stale)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build
        build.build()
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.tcr_last_error.restype = C.c_char_p
    lib.tcr_create.argtypes = [C.c_int, C.POINTER(TcrParams), C.POINTER(vp)]
    lib.tcr_destroy.argtypes = [vp]
    lib.tcr_set_stream.argtypes = [vp, vp]
    lib.tcr_synchronize.argtypes = [vp]
    lib.tcr_upload_static.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp]
    lib.tcr_upload_masks.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
    lib.tcr_alloc_tables.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp]
    lib.tcr_upload_month.argtypes = [vp, C.c_int, C.POINTER(vp)]
    lib.tcr_upload_month_dev.argtypes = [vp, C.c_int, vp]
    lib.tcr_upload_months.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.tcr_env_interp.argtypes = [vp, C.c_int64, vp, vp, vp, vp, C.c_int]
    lib.tcr_integrate.argtypes = [vp, C.c_int64] + [vp] * 14 + [C.c_int]
    lib.tcr_run_years.argtypes = [vp, C.c_int, vp, vp, C.c_uint32, C.c_int] + [vp] * 9 + [C.POINTER(TcrYearStats), C.c_int]
    lib.tcr_seed_attempts.argtypes = [vp, C.c_int, C.c_int32, C.c_uint32, C.c_int64, C.c_int64] + [vp] * 8
    lib.tcr_set_tuning.argtypes = [vp, C.c_int, C.c_int64, C.c_int64, C.c_int]
    lib.tcr_set_workspace_budget.argtypes = [vp, C.c_double, C.c_int64]
    lib.tcr_fourier_ring_nodes.argtypes = [vp]
    lib.tcr_launch_count.argtypes = [vp]
    lib.tcr_launch_count.restype = C.c_int64
    lib.tcr_set_interp_variant.argtypes = [vp, C.c_int]
    lib.tcr_set_timing.argtypes = [vp, C.c_int]
    lib.tcr_kernel_time.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.tcr_poi_vmax.argtypes = [vp, C.c_int64, C.c_int, vp, vp, vp, C.c_double, C.c_double, C.c_double, C.c_double, vp, C.c_int]
    lib.tcr_exceedance.argtypes = [vp, C.c_int64, vp, C.c_int, vp, vp, C.c_int]
    lib.tcr_prepare_month.argtypes = [vp, C.c_int, vp] + [vp] * 9
    lib.tcr_wind_stats.argtypes = [vp, C.c_int, C.c_int64, C.c_int64, vp, vp, vp, vp, C.c_int, vp, vp, C.c_int]
    lib.tcr_set_entropy_table.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
    lib.tcr_thermo_month.argtypes = [vp, C.c_int64, C.c_int, vp, vp, vp, vp, vp, C.c_double, C.c_int, vp, vp, vp, C.c_int]
    lib.tcr_set_entropy_table_reversible.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]
    lib.tcr_thermo_month_reversible.argtypes = lib.tcr_thermo_month.argtypes
    lib.tcr_rhs_eval.argtypes = [vp, C.c_int64] + [vp] * 7
    lib.tcr_set_shard.argtypes = [vp, C.c_int, C.c_int, ALLREDUCE_FN, vp]
    lib.tcr_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    lib.tcr_host_free.argtypes = [vp]
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise TcrError(load().tcr_last_error().decode("utf-8", "replace"))
