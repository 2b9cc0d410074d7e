"""ctypes binding of libnof_sm100.so (C ABI declared in include/nof.h).

The product path has NO fallback: if the library is missing it is built in-tree with nvcc; if that fails, or a
function is called without a CUDA device, an exception is raised.
"""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('NOF_LIB', os.path.join(HERE, 'lib', 'libnof_sm100.so'))   # NOF_LIB: load a tuning variant

NOF_F32, NOF_F16 = 0, 1
_vp, _i32, _u32, _u64, _f32, _sz, _i64 = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64, C.c_float, C.c_size_t, C.c_int64


class NofMarchCfg(C.Structure):
    _fields_ = [('N', _i32), ('ray_dim', _i32), ('S_occ', _i32), ('S_depth', _i32), ('level', _i32), ('I_max', _i32),
                ('trunc', _f32), ('near_sc', _f32), ('far_sc', _f32), ('neg_trunc_ratio', _f32), ('perturb', _i32),
                ('seed', _u64), ('offset', _u64), ('offset_ptr', _vp), ('trunc_ptr', _vp)]


class NofPrologue(C.Structure):
    _fields_ = [('pool', _vp), ('ids', _vp), ('n_ids', _i64), ('batch', _vp), ('N', _i32), ('ray_dim', _i32), ('cursor', _vp),
                ('pose_data', _vp), ('c2w', _vp), ('tf', _vp), ('F', _i32), ('max_trans', _f32), ('max_rot_deg', _f32),
                ('tick', _vp), ('done', _vp), ('trunc_table', _vp), ('trunc_len', _i32), ('gstep', _vp), ('trunc_out', _vp)]


class NofStep(C.Structure):
    _fields_ = [('N', _i32), ('S', _i32), ('L', _i32), ('C', _i32), ('F', _i32), ('ff', _i32), ('ray_dim', _i32), ('amp', _i32),
                ('S_log2', _f32), ('H', _i32), ('offsets', _vp), ('table_f32', _vp), ('table_f16', _vp),
                ('mlp', _vp), ('feat', _vp), ('rays', _vp), ('tf', _vp), ('z_vals', _vp),
                ('trunc', _f32), ('near_sc', _f32), ('far_sc', _f32), ('sdf_lambda', _f32), ('neg_trunc_ratio', _f32),
                ('rgb_weight', _f32), ('fs_weight', _f32), ('empty_weight', _f32), ('trunc_weight', _f32), ('fs_sdf', _f32),
                ('fs_rgb_weight', _f32), ('first_frame_weight', _f32),
                ('loss_scale', _vp), ('need_pose_grad', _i32),
                ('grad_table', _vp), ('grad_mlp', _vp), ('grad_tf', _vp), ('grad_feat', _vp), ('losses', _vp), ('found_inf', _vp),
                ('rgb_map', _vp), ('raw', _vp), ('valid_samples', _vp), ('weights', _vp), ('workspace', _vp), ('eikonal_weight', _f32), ('trunc_ptr', _vp)]


class NofAdamSeg(C.Structure):
    _fields_ = [('param', _vp), ('grad', _vp), ('exp_avg', _vp), ('exp_avg_sq', _vp), ('shadow_f16', _vp), ('n', _sz), ('lr', _f32), ('lr_ptr', _vp)]


_SIGS = {
    'nof_version': (C.c_int, []),
    'nof_last_error': (C.c_char_p, []),
    'nof_device_info': (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'nof_grid_encode_forward': (C.c_int, [_vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _f32, _u32, C.c_int, _vp, _u32, C.c_int, C.c_int, _vp]),
    'nof_grid_encode_backward': (C.c_int, [_vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _f32, _u32, C.c_int, _vp, _vp, _u32, C.c_int, C.c_int, _vp]),
    'nof_grid_level_scales': (C.c_int, [_f32, _u32, C.c_int, _vp, _vp]),
    'nof_sample_rays_uniform_occupied_voxels': (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    'nof_postprocess_octree_ray_tracing': (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    'nof_gather_rays': (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    'nof_pose_forward': (C.c_int, [_vp, _vp, _vp, C.c_int, _f32, _f32, _vp]),
    'nof_step_prologue': (C.c_int, [C.POINTER(NofPrologue), _vp]),
    'nof_pose_backward': (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _f32, _f32, _vp, _vp]),
    'nof_ray_march': (C.c_int, [C.POINTER(NofMarchCfg), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'nof_mlp_param_count': (_sz, [C.c_int, C.c_int]),
    'nof_mlp_param_offsets': (C.c_int, [C.c_int, C.c_int, C.POINTER(_i32)]),
    'nof_step_workspace_bytes': (_sz, [C.POINTER(NofStep)]),
    'nof_step_fused': (C.c_int, [C.POINTER(NofStep), _vp]),
    'nof_set_amp_impl': (C.c_int, [C.c_int]),
    'nof_adam_step': (C.c_int, [C.POINTER(NofAdamSeg), C.c_int, _f32, _f32, _f32, _vp, _vp, _vp, _vp, _vp]),
    'nof_adam_update': (C.c_int, [C.POINTER(NofAdamSeg), C.c_int, _f32, _f32, _f32, _vp, _vp, _vp, _vp]),
    'nof_adam_finish': (C.c_int, [_vp, _vp, _vp, _vp, _f32, _f32, _vp]),
    'nof_adam_tile_count': (C.c_int, [C.POINTER(NofAdamSeg), C.c_int]),
    'nof_adam_update_shared': (C.c_int, [C.POINTER(NofAdamSeg), C.c_int, _f32, _f32, _f32, _vp, _vp, _vp, _vp, C.c_int, _vp]),
    'nof_query_sdf': (C.c_int, [C.POINTER(NofStep), _vp, _vp, _i64, _vp]),
    'nof_cloud_within_radius': (C.c_int, [_vp, _i64, _vp, _vp, _f32, _f32, C.c_int, _f32, _vp, _vp]),
    'nof_marching_tets_count': (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _f32, _vp, _vp]),
    'nof_marching_tets_emit': (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _f32, _vp, _vp, _vp, _vp]),
}
EXPORTS = tuple(_SIGS)
_lib = None


class NofError(RuntimeError):
    pass


def load(build_if_missing=True):
    """Load (building first if absent) the shared library; raises if it cannot be had. No CPU fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise NofError(f'{LIB_PATH} is missing; run `python -m bundlesdf_b200.build`')
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)           # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=''):
    if rc != 0:
        msg = load().nof_last_error().decode('utf-8', 'replace')
        raise NofError(f'{what} failed (code {rc}): {msg}')


def ptr(t, pinned_ok=False):
    """Device pointer of a tensor (None -> NULL). The tensor must be CUDA + contiguous; with pinned_ok a PINNED host tensor is accepted
    too (page-locked memory is device-addressable under unified addressing: the kernel reads / writes it across PCIe)."""
    if t is None:
        return None
    if pinned_ok and not t.is_cuda and t.is_pinned() and t.is_contiguous():
        return t.data_ptr()
    if not t.is_cuda:
        raise NofError('libnof_sm100 needs CUDA tensors: there is no CPU fallback for the Neural-Object-Field hot path')
    if not t.is_contiguous():
        raise NofError('tensor must be contiguous')
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def require_cuda():
    if not torch.cuda.is_available():
        raise NofError('bundlesdf_b200 requires a CUDA device (sm_100a); no CPU fallback exists for the hot path')
