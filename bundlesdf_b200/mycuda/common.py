"""`common` module of the reference (mycuda/bindings.cpp:15-19, common.h:28-29): the two hot-path functions with the
reference's signatures. rayColorToTextureImageCUDA (offline texture baking) is out of scope."""
import torch

from .. import _lib


def _check_input(t, name):
    if not t.is_cuda:
        raise RuntimeError(f'{name} must be a CUDA tensor')            # common.h:19 CHECK_CUDA
    if not t.is_contiguous():
        raise RuntimeError(f'{name} must be contiguous')               # common.h:20 CHECK_CONTIGUOUS


def sampleRaysUniformOccupiedVoxels(z_in_out, z_sampled, z_vals):
    """common.cu:107-125. Returns z_vals (filled in place). Raises instead of hanging the GPU when the interval list
    and the cumulative sample disagree (the reference kernel prints and spins forever, common.cu:66-71)."""
    lib = _lib.load()
    for t, n in ((z_in_out, 'z_in_out'), (z_sampled, 'z_sampled'), (z_vals, 'z_vals')):
        _check_input(t, n)
    if z_vals.shape != z_sampled.shape:
        raise RuntimeError('z_vals.sizes()==z_sampled.sizes()')        # common.cu:112 AT_ASSERTM
    if z_in_out.dtype != torch.float32:
        raise RuntimeError('only float32 is built (the reference dispatches float/double)')
    N, S = z_sampled.shape
    I = z_in_out.shape[1]
    err = torch.zeros(1, dtype=torch.int32, device=z_vals.device)
    _lib.check(lib.nof_sample_rays_uniform_occupied_voxels(z_in_out.data_ptr(), z_sampled.data_ptr(), z_vals.data_ptr(),
                                                           N, I, S, err.data_ptr(), _lib.stream()), 'sampleRaysUniformOccupiedVoxels')
    if int(err.item()) != 0:            # one host sync, like the reference's implicit one on its legacy-stream launch + later .item()s
        raise RuntimeError('sampleRaysUniformOccupiedVoxels: a cumulative sample ran past the last interval of its ray '
                           '(the reference kernel prints "z_remain is invalid" and never returns, common.cu:66-71,87-92); '
                           'the affected samples were clamped to the end of the last interval')
    return z_vals


def postprocessOctreeRayTracing(ray_index, depth_in_out, unique_intersect_ray_ids, start_poss, max_intersections, N_rays):
    """common.cu:151-167. Allocates and returns the padded [N_rays, max_intersections, 2] tensor on the device of the
    inputs (the reference hard-codes cuda:0)."""
    lib = _lib.load()
    for t, n in ((ray_index, 'ray_index'), (depth_in_out, 'depth_in_out'), (start_poss, 'start_poss')):
        _check_input(t, n)
    out = torch.empty((N_rays, max_intersections, 2), dtype=torch.float32, device=depth_in_out.device)
    _lib.check(lib.nof_postprocess_octree_ray_tracing(ray_index.data_ptr(), depth_in_out.data_ptr(),
                                                      unique_intersect_ray_ids.data_ptr(), start_poss.data_ptr(),
                                                      ray_index.shape[0], unique_intersect_ray_ids.shape[0],
                                                      int(max_intersections), int(N_rays), out.data_ptr(), _lib.stream()),
               'postprocessOctreeRayTracing')
    return out
