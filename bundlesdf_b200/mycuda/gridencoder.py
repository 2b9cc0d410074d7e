"""`gridencoder` module of the reference (mycuda/torch_ngp_grid_encoder/bindings.cpp:16-19), same two functions and
argument order (gridencoder.h:23-24), dispatched to nof_grid_encode_forward/backward on torch's CURRENT stream."""
import torch

from .. import _lib


def _dtype_id(t):
    if t.dtype == torch.float32:
        return _lib.NOF_F32
    if t.dtype == torch.float16:
        return _lib.NOF_F16
    raise RuntimeError(f'embeddings must be float32 or float16, got {t.dtype}')   # reference: AT_DISPATCH error


def _check(t, name, dtype=None):
    if not t.is_cuda:
        raise RuntimeError(f'{name} must be a CUDA tensor')            # gridencoder.cu:27 CHECK_CUDA
    if not t.is_contiguous():
        raise RuntimeError(f'{name} must be a contiguous tensor')      # :28 CHECK_CONTIGUOUS
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f'{name} must be a {dtype} tensor')


def grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, calc_grad_inputs, dy_dx, gridtype, align_corners):
    lib = _lib.load()
    _check(inputs, 'inputs', torch.float32); _check(embeddings, 'embeddings'); _check(offsets, 'offsets', torch.int32)
    _check(outputs, 'outputs', embeddings.dtype); _check(dy_dx, 'dy_dx', embeddings.dtype)
    _lib.check(lib.nof_grid_encode_forward(inputs.data_ptr(), embeddings.data_ptr(), offsets.data_ptr(), outputs.data_ptr(),
                                           int(B), int(D), int(C), int(L), float(S), int(H), int(bool(calc_grad_inputs)),
                                           dy_dx.data_ptr(), int(gridtype), int(bool(align_corners)), _dtype_id(embeddings),
                                           _lib.stream()), 'grid_encode_forward')


def grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, calc_grad_inputs, dy_dx,
                         grad_inputs, gridtype, align_corners):
    lib = _lib.load()
    _check(grad, 'grad'); _check(inputs, 'inputs', torch.float32); _check(embeddings, 'embeddings'); _check(offsets, 'offsets', torch.int32)
    _check(grad_embeddings, 'grad_embeddings', grad.dtype); _check(dy_dx, 'dy_dx'); _check(grad_inputs, 'grad_inputs')
    _lib.check(lib.nof_grid_encode_backward(grad.data_ptr(), inputs.data_ptr(), embeddings.data_ptr(), offsets.data_ptr(),
                                            grad_embeddings.data_ptr(), int(B), int(D), int(C), int(L), float(S), int(H),
                                            int(bool(calc_grad_inputs)), dy_dx.data_ptr(), grad_inputs.data_ptr(), int(gridtype),
                                            int(bool(align_corners)), _dtype_id(grad), _lib.stream()), 'grid_encode_backward')
