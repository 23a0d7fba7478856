"""ctypes binding of the C ABI declared in ``include/jaxpm_b200.h``.

There is no CPU fallback: if ``libjaxpm_b200.so`` is missing (run
``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C jaxpm_b200/csrc``)
or a tensor is not a CUDA tensor, the call fails loudly.

PyTorch is used only as the device-memory / stream / process-group provider:
every compute op on the path goes through the symbols bound here.
"""
import ctypes as C
import os

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libjaxpm_b200.so")
_lib = None

vp, f32, i32, i64 = C.c_void_p, C.c_float, C.c_int32, C.c_int64

# name -> argtypes; every symbol of include/jaxpm_b200.h (tests check this list against the header)
SIGNATURES = {
    "jpm_abi_version": ([], i32),
    "jpm_last_error_string": ([], C.c_char_p),
    "jpm_device_info": ([C.c_char_p, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)], i32),
    "jpm_cic_paint_f32": ([vp, vp, vp, vp, f32, i64, i32, i32, i32, i32, i32, i32], i32),
    "jpm_cic_paint_dx_f32": ([vp, vp, vp, vp, f32, i32, i32, i32, i32, i32], i32),
    "jpm_cic_paintgrad_f32": ([vp, vp, vp, vp, vp, f32, i64, i32, i32, i32, i32, i32, i32], i32),
    "jpm_cic_paint_2d_f32": ([vp, vp, vp, vp, i64, i32, i32], i32),
    "jpm_density_plane_f32": ([vp, vp, vp, i64, f32, C.c_double, C.c_double, i32], i32),
    "jpm_cic_cell_index_i32": ([vp, vp, vp, i64, i32, i32, i32, i32, i32, i32], i32),
    "jpm_cic_read_f32": ([vp, vp, vp, vp, i64, i32, i32, i32], i32),
    "jpm_cic_read_dx_f32": ([vp, vp, vp, vp, i32, i32, i32, i32, i32], i32),
    "jpm_cic_read3_f32": ([vp, vp, vp, vp, vp, vp, f32, i64, i32, i32, i32, i32, i32, i32], i32),
    "jpm_cic_read3_kick_drift_f32": ([vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, f32, f32, i32,
                                      i64, i32, i32, i32, i32, i32, i32], i32),
    "jpm_cic_readgrad_f32": ([vp, vp, vp, vp, vp, vp, f32, i64, i32, i32, i32, i32, i32, i32], i32),
    "jpm_plan_create": ([C.POINTER(vp), i32, i32, i32], i32),
    "jpm_plan_destroy": ([vp], i32),
    "jpm_fft3d_r2c": ([vp, vp, vp, vp], i32),
    "jpm_ifft3d_c2r": ([vp, vp, vp, vp, i32], i32),
    "jpm_greens_grad_c64": ([vp, vp, vp, vp, f32, f32, vp, i32, f32], i32),
    "jpm_greens_div_c64": ([vp, vp, vp, vp, f32, f32, vp, i32, f32], i32),
    "jpm_lpt2_shear_c64": ([vp, vp, vp, vp, f32], i32),
    "jpm_lpt2_source_f32": ([vp, vp, vp, i64], i32),
    "jpm_lpt2_source_adj_f32": ([vp, vp, vp, vp, i64], i32),
    "jpm_lpt2_shear_adj_c64": ([vp, vp, vp, vp, f32], i32),
    "jpm_kfilter_logtab_c64": ([vp, vp, vp, vp, vp, i32, f32, f32, f32, f32, f32, f32], i32),
    "jpm_density_to_force_meshes": ([vp, vp, vp, vp, f32, vp, i32, f32], i32),
    "jpm_density_to_force_meshes_fused": ([vp, vp, vp, vp, f32, vp, i32, f32], i32),
    "jpm_density_to_potential_fused": ([vp, vp, vp, vp, f32, vp, i32, f32], i32),
    "jpm_pm_step_f32": ([vp, vp, vp, vp, f32, f32, i32], i32),
    "jpm_pm_step_host_f32": ([vp, vp, vp, vp, vp, vp, f32, f32, i32], i32),
    "jpm_pk_bin_c64": ([vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, i32, C.POINTER(i32), i32, C.POINTER(f32), f32, i32,
                        vp], i32),
    "jpm_pk_weight_c64": ([vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, i32, C.POINTER(i32), i32, C.POINTER(f32), vp,
                           f32], i32),
    "jpm_kseparable_c64": ([vp, vp, vp, vp, vp, vp, i32, i32, i32, f32], i32),
    "jpm_plan_padded_get_f32": ([vp, vp, i32, vp, C.POINTER(i32)], i32),
    "jpm_slab_create": ([C.POINTER(vp), i32, i32, i32, i32, i32, i32], i32),
    "jpm_slab_create_ex": ([C.POINTER(vp), i32, i32, i32, i32, i32, i32, i32, i32], i32),
    "jpm_slab_ipc_handle": ([vp, vp, i32], i32),
    "jpm_slab_attach_ipc": ([vp, vp, i32], i32),
    "jpm_slab_attach_ptrs": ([vp, C.POINTER(vp), i32], i32),
    "jpm_slab_base": ([vp, C.POINTER(vp), C.POINTER(i64)], i32),
    "jpm_slab_get_interior_f32": ([vp, vp, i32, vp], i32),
    "jpm_slab_set_density_f32": ([vp, vp, vp], i32),
    "jpm_slab_forces": ([vp, vp, f32], i32),
    "jpm_slab_check": ([vp, vp], i32),
    "jpm_slab_ghost_width": ([vp, vp, C.POINTER(i32)], i32),
    "jpm_slab_halo_exceeded": ([vp, vp, C.POINTER(i32)], i32),
    "jpm_sim_create": ([C.POINTER(vp), vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32], i32),
    "jpm_sim_create_ex": ([C.POINTER(vp), vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32], i32),
    "jpm_sim_destroy": ([vp], i32),
    "jpm_sim_load": ([vp, vp, vp, vp], i32),
    "jpm_sim_store": ([vp, vp, vp, vp], i32),
    "jpm_sim_paint": ([vp, vp, vp], i32),
    "jpm_sim_read_kick_drift": ([vp, vp, vp, vp, vp, f32, f32], i32),
    "jpm_sim_forces": ([vp, vp, vp, f32, f32, vp, i32, f32], i32),
    "jpm_sim_forces_batched": ([vp, vp, vp, vp, i32, f32, f32, vp, i32, f32], i32),
    "jpm_sim_step": ([vp, vp, f32, f32], i32),
    "jpm_sim_step_host_f32": ([vp, vp, vp, vp, vp, vp, f32, f32], i32),
    "jpm_sim_steps_host_f32": ([vp, vp, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(f32), C.POINTER(f32)], i32),
    "jpm_sim_step_profile": ([vp, vp, f32, f32, C.POINTER(C.c_char_p), C.POINTER(f32), i32, C.POINTER(i32)], i32),
    "jpm_sim_stats_host": ([vp, vp, C.POINTER(i64)], i32),
    "jpm_sim_set_force_mode": ([vp, i32], i32),
    "jpm_sim_force_info": ([vp, vp, C.POINTER(C.c_double)], i32),
    "jpm_cic_readgrad3_f32": ([vp, vp, vp, vp, vp, vp, vp, f32, i64, i32, i32, i32, i32, i32, i32, i32], i32),
    "jpm_cic_paint3_f32": ([vp, vp, vp, vp, f32, i64, i32, i32, i32, i32, i32, i32], i32),
    "jpm_fd_divergence3_f32": ([vp, vp, vp, i32, i32, i32], i32),
    "jpm_cic_paint_f64": ([vp, vp, vp, vp, C.c_double, i64, i32, i32, i32], i32),
    "jpm_cic_paint_dx_f64": ([vp, vp, vp, vp, C.c_double, i32, i32, i32, i32, i32], i32),
    "jpm_cic_read_f64": ([vp, vp, vp, vp, i64, i32, i32, i32], i32),
    "jpm_cic_read_dx_f64": ([vp, vp, vp, vp, i32, i32, i32, i32, i32], i32),
    "jpm_kernel_launch_count": ([], i64),
    "jpm_normal_field_f32": ([vp, vp, i32, i32, i32, i32, i32, i32, C.c_uint64, C.c_uint32], i32),
    "jpm_linear_field_f32": ([vp, vp, vp, vp, vp, i32, f32, f32, f32, f32, f32, f32], i32),
    "jpm_axpby_f32": ([vp, vp, f32, vp, f32, vp, i64], i32),
    "jpm_grid_plus_disp_f32": ([vp, vp, vp, i32, i32, i32, i32, i32], i32),
    "jpm_fft1d_create": ([C.POINTER(vp), i32, i64, i32], i32),
    "jpm_fft1d_destroy": ([vp], i32),
    "jpm_fft1d_exec": ([vp, vp, vp, vp, i32], i32),
    "jpm_transpose_c64": ([vp, vp, vp, i32, i32, i64, i64, i64, i64, i64], i32),
    "jpm_copy2d_c64": ([vp, vp, vp, i64, i32, i64, i64], i32),
    "jpm_kspace_local_c64": ([vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, f32, f32,
                              vp, i32, f32], i32),
    "jpm_pack_box_f32": ([vp, vp, vp, i32, i32, i32, i32, i32, i32], i32),
    "jpm_unpack_box_f32": ([vp, vp, vp, i32, i32, i32, i32, i32, i32, i32], i32),
}


class JpmError(RuntimeError):
    pass


def lib_path():
    return _LIB_PATH


def load():
    """Load the shared library (once) and bind every symbol."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError(
                f"{_LIB_PATH} not built: run `make -C jaxpm_b200/csrc` (or __graft_entry__.build()). "
                "jaxpm_b200 has no CPU fallback.")
        lib = C.CDLL(_LIB_PATH)
        for name, (argtypes, restype) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so is stale
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = lib
    return _lib


def call(name, *args):
    """Call an int32-returning entry point and raise on a nonzero status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise JpmError(f"{name} -> {rc}: {lib.jpm_last_error_string().decode()}")


def launch_count():
    return int(load().jpm_kernel_launch_count())


# ---- tensor plumbing -----------------------------------------------------------
def _torch():
    import torch
    return torch


def ptr(t, dtype=None):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    torch = _torch()
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise JpmError("jaxpm_b200 ops need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise JpmError("tensor must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise JpmError(f"expected dtype {dtype}, got {t.dtype}")
    return t.data_ptr()


def stream():
    return _torch().cuda.current_stream().cuda_stream


def as_f32(x, device=None):
    """Contiguous float32 CUDA tensor view/copy of ``x``."""
    torch = _torch()
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x, dtype=torch.float32)
    if device is not None and x.device != device:
        x = x.to(device)
    if not x.is_cuda:
        raise JpmError("jaxpm_b200 ops need CUDA tensors (no CPU fallback)")
    if x.dtype != torch.float32:
        x = x.to(torch.float32)
    return x.contiguous()
