"""GPU parity tests for the op-level C-ABI entry points (through bundlesdf_b200.mycuda / ops) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import nof_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _table(L, finest, log2T, seed, scale=0.5):
    offsets, pls = O.grid_offsets(L, 16, finest, log2T)
    g = torch.Generator().manual_seed(seed)
    emb = (torch.rand(int(offsets[-1]), 2, generator=g) * 2 - 1) * scale
    return offsets, float(np.log2(pls)), emb


def _points(B, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, generator=g)
    x[0] = torch.tensor([0.0, 0.0, 0.0]); x[1] = torch.tensor([1.0, 1.0, 1.0]); x[2] = torch.tensor([0.5, 1.0, 0.0])
    x[3] = torch.tensor([1.0001, 0.5, 0.5]); x[4] = torch.tensor([0.5, -1e-6, 0.5])          # out of bounds -> zeros
    return x


@pytest.mark.parametrize('L,finest,log2T', [(4, 128, 22), (16, 256, 19), (16, 512, 16)])
def test_grid_forward_fp32_vs_oracle(L, finest, log2T):
    from bundlesdf_b200.mycuda import gridencoder
    offsets, S, emb = _table(L, finest, log2T, 1)
    B = 2053
    x = _points(B, 2)
    from bundlesdf_b200 import ops
    scales = ops.grid_level_scales(S, 16, L).cpu().numpy()          # CUDA exp2f, not libm (differs by an ulp at some levels)
    out_o, dy_o = O.grid_encode(x, emb, offsets, S, 16, exact_fma=True, want_dydx=True, scales=scales)
    xd, ed, od = x.to(DEV), emb.to(DEV), torch.from_numpy(offsets).to(DEV)
    out = torch.empty(L, B, 2, device=DEV)
    dy = torch.empty(B, L * 3 * 2, device=DEV)
    gridencoder.grid_encode_forward(xd, ed, od, out, B, 3, 2, L, S, 16, True, dy, 0, False)
    got = out.permute(1, 0, 2).reshape(B, L * 2).cpu()
    np.testing.assert_allclose(got.numpy(), out_o.numpy(), rtol=1e-5, atol=1e-7)
    exact = (got == out_o).float().mean().item()
    assert exact == 1.0, f'only {exact:.4f} of the fp32 forward values are bit-identical to the FMA-emulating oracle'
    np.testing.assert_allclose(dy.view(B, L, 3, 2).cpu().numpy(), dy_o.numpy(), rtol=2e-4, atol=2e-4)
    assert torch.all(got[3] == 0) and torch.all(got[4] == 0)


def test_grid_forward_fp16_vs_oracle():
    from bundlesdf_b200.mycuda import gridencoder
    L, finest, log2T = 16, 256, 19
    offsets, S, emb = _table(L, finest, log2T, 3)
    B = 1024
    x = _points(B, 4)
    emb_h = emb.half()
    out_o = O.grid_encode(x, emb_h.float(), offsets, S, 16, exact_fma=False)
    out = torch.empty(L, B, 2, device=DEV, dtype=torch.half)
    dy = torch.empty(1, device=DEV, dtype=torch.half)
    gridencoder.grid_encode_forward(x.to(DEV), emb_h.to(DEV), torch.from_numpy(offsets).to(DEV), out, B, 3, 2, L, S, 16, False, dy, 0, False)
    got = out.permute(1, 0, 2).reshape(B, L * 2).float().cpu()
    np.testing.assert_allclose(got.numpy(), out_o.numpy(), rtol=0, atol=3e-3)          # fp16 accumulate: a few ulp(0.5)


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
def test_grid_backward_vs_oracle_autograd(dtype):
    from bundlesdf_b200.mycuda import gridencoder
    L, finest, log2T = 16, 256, 14
    offsets, S, emb = _table(L, finest, log2T, 5)
    B = 777
    x = _points(B, 6).requires_grad_(True)
    embr = (emb.to(dtype).float()).requires_grad_(True)
    out_o = O.grid_encode(x, embr, offsets, S, 16, exact_fma=False)
    g = torch.Generator().manual_seed(7)
    gy = torch.randn(B, L * 2, generator=g) * 0.1
    gy = gy.to(dtype).float()
    (out_o * gy).sum().backward()
    xd, od = x.detach().to(DEV), torch.from_numpy(offsets).to(DEV)
    ed = emb.to(dtype).to(DEV)
    out = torch.empty(L, B, 2, device=DEV, dtype=dtype)
    dy = torch.empty(B, L * 3 * 2, device=DEV, dtype=dtype)
    gridencoder.grid_encode_forward(xd, ed, od, out, B, 3, 2, L, S, 16, True, dy, 0, False)
    grad = gy.view(B, L, 2).permute(1, 0, 2).contiguous().to(dtype).to(DEV)
    gemb = torch.zeros_like(ed)
    gin = torch.zeros(B, 3, device=DEV, dtype=dtype)
    gridencoder.grid_encode_backward(grad, xd, ed, od, gemb, B, 3, 2, L, S, 16, True, dy, gin, 0, False)
    tol = dict(rtol=1e-4, atol=1e-5) if dtype == torch.float32 else dict(rtol=3e-2, atol=3e-2)
    np.testing.assert_allclose(gemb.float().cpu().numpy(), embr.grad.numpy(), **tol)
    tol_in = dict(rtol=1e-3, atol=1e-3) if dtype == torch.float32 else dict(rtol=5e-2, atol=0.5)
    np.testing.assert_allclose(gin.float().cpu().numpy(), x.grad.numpy(), **tol_in)


def test_grid_encoder_module_autograd_matches_oracle():
    """The nn.Module mirror (GridEncoder) end to end, fp32, incl. the (x+1)/2 mapping and the [L,B,C]->[B,L*C] permute."""
    from bundlesdf_b200.mycuda.torch_ngp_grid_encoder.grid import GridEncoder
    torch.manual_seed(0)
    enc = GridEncoder(n_levels=4, desired_resolution=128, log2_hashmap_size=22).to(DEV)
    with torch.no_grad():
        enc.embeddings.uniform_(-0.5, 0.5)
    x = (torch.rand(300, 3) * 2 - 1).to(DEV).requires_grad_(True)
    y = enc(x)
    gy = torch.randn_like(y)
    (y * gy).sum().backward()
    xo = x.detach().cpu().requires_grad_(True)
    eo = enc.embeddings.detach().cpu().requires_grad_(True)
    yo = O.grid_encode((xo + 1) / 2, eo, enc.offsets.cpu().numpy(), float(np.log2(enc.per_level_scale)), 16, exact_fma=False)
    (yo * gy.cpu()).sum().backward()
    np.testing.assert_allclose(y.detach().cpu().numpy(), yo.detach().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(enc.embeddings.grad.cpu().numpy(), eo.grad.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(x.grad.cpu().numpy(), xo.grad.numpy(), rtol=1e-3, atol=1e-3)


def _random_intervals(N, I, seed):
    rng = np.random.default_rng(seed)
    io = np.zeros((N, I, 2), np.float32)
    for r in range(N):
        k = rng.integers(0, I + 1)
        t = rng.uniform(0.5, 2.0)
        for j in range(k):
            a = t + rng.uniform(0.0, 0.2) * (rng.random() < 0.5)
            b = a + rng.uniform(0.01, 0.3)
            io[r, j] = (a, b)
            t = b
    return io


def test_interval_walk_bit_exact_vs_oracle():
    from bundlesdf_b200.mycuda import common
    N, I, S = 301, 7, 64
    io = _random_intervals(N, I, 0)
    rng = np.random.default_rng(1)
    total = (io[:, :, 1] - io[:, :, 0]).sum(-1, dtype=np.float32)
    zs = (rng.random((N, S), dtype=np.float32) * total[:, None] * 0.999).astype(np.float32)
    want, err = O.interval_walk(io, zs)
    assert not err
    zv = torch.zeros(N, S, device=DEV)
    got = common.sampleRaysUniformOccupiedVoxels(torch.from_numpy(io).to(DEV), torch.from_numpy(zs).to(DEV), zv)
    np.testing.assert_array_equal(got.cpu().numpy(), want)


def test_interval_walk_flags_instead_of_hanging():
    """Where the reference kernel prints and spins forever (common.cu:66-71) we clamp and raise a flag."""
    from bundlesdf_b200 import _lib
    lib = _lib.load()
    io = torch.tensor([[[1.0, 1.5], [0.0, 0.0]]], device=DEV)
    zs = torch.tensor([[0.2, 0.9]], device=DEV)          # 0.9 > total length 0.5
    zv = torch.zeros(1, 2, device=DEV)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    _lib.check(lib.nof_sample_rays_uniform_occupied_voxels(io.data_ptr(), zs.data_ptr(), zv.data_ptr(), 1, 2, 2, flag.data_ptr(), _lib.stream()))
    torch.cuda.synchronize()
    assert flag.item() == 1
    np.testing.assert_allclose(zv.cpu().numpy(), [[1.2, 1.5]])


def test_postprocess_bit_exact_vs_oracle():
    from bundlesdf_b200.mycuda import common
    rng = np.random.default_rng(3)
    n_rays = 50
    ray_index, dio = [], []
    for r in range(n_rays):
        if rng.random() < 0.2:
            continue
        k = rng.integers(1, 9)
        t = rng.uniform(0.5, 1.0)
        for j in range(k):
            a = t
            b = a + (rng.uniform(0.0, 0.3) if rng.random() > 0.2 else 5e-5)
            if rng.random() < 0.05:
                a, b = b, a
            ray_index.append(r); dio.append((a, b)); t = b
    ray_index = np.array(ray_index, np.int64); dio = np.array(dio, np.float32)
    uniq, counts = np.unique(ray_index, return_counts=True)
    start = np.concatenate([[0], np.cumsum(counts[:-1])]).astype(np.int64)
    mi = int(counts.max())
    want = O.postprocess_octree_ray_tracing(ray_index, dio, uniq, start, mi, n_rays)
    got = common.postprocessOctreeRayTracing(torch.from_numpy(ray_index).to(DEV), torch.from_numpy(dio).to(DEV),
                                             torch.from_numpy(uniq).to(DEV), torch.from_numpy(start).to(DEV), mi, n_rays)
    np.testing.assert_array_equal(got.cpu().numpy(), want)


def test_pose_forward_backward_vs_oracle():
    from bundlesdf_b200 import ops
    g = torch.Generator().manual_seed(0)
    F = 9
    data = torch.randn(F, 6, generator=g) * 0.4
    data[2, 3:] = 0.0
    data[3, 3:] *= 1e-3
    c2w = torch.eye(4).repeat(F, 1, 1)
    c2w[:, :3, :] = torch.randn(F, 3, 4, generator=g)
    mt, mr = 0.02 * 5.0, 20.0
    dref = data.clone().requires_grad_(True)
    T = O.pose_matrices(dref, mt, mr) @ c2w
    gt = torch.randn(F, 12, generator=g)
    (T[:, :3, :].reshape(F, 12) * gt).sum().backward()
    tf = ops.pose_forward(data.to(DEV), c2w.to(DEV), mt, mr)
    np.testing.assert_allclose(tf.cpu().numpy(), T[:, :3, :].reshape(F, 12).detach().numpy(), rtol=1e-5, atol=1e-6)
    gp = torch.zeros(F, 6, device=DEV)
    ops.pose_backward(data.to(DEV), c2w.to(DEV), gt.to(DEV), gp, mt, mr)
    np.testing.assert_allclose(gp.cpu().numpy(), dref.grad.numpy(), rtol=1e-4, atol=1e-5)
    assert torch.all(gp[0] == 0)


def test_adam_matches_torch_adam():
    from bundlesdf_b200 import ops
    g = torch.Generator().manual_seed(0)
    sizes = [100003, 9108, 24]
    ps = [torch.randn(n, generator=g) for n in sizes]
    ref = [p.clone().requires_grad_(True) for p in ps]
    opt = torch.optim.Adam([{'params': ref[:2], 'lr': 0.01}, {'params': ref[2:], 'lr': 0.003}], betas=(0.9, 0.999), eps=1e-15)
    dev = [dict(param=p.to(DEV), grad=torch.zeros(n, device=DEV), exp_avg=torch.zeros(n, device=DEV), exp_avg_sq=torch.zeros(n, device=DEV),
                shadow_f16=(torch.zeros(n, device=DEV, dtype=torch.half) if i == 0 else None), lr=(0.01 if i < 2 else 0.003))
           for i, (p, n) in enumerate(zip(ps, sizes))]
    step = torch.zeros(8, dtype=torch.int32, device=DEV)      # [0] count, [1..7] library scratch (include/nof.h)
    for it in range(5):
        grads = [torch.randn(n, generator=g) * (0.0 if (it == 2 and i == 1) else 1.0) for i, n in enumerate(sizes)]
        for r, gr in zip(ref, grads):
            r.grad = gr.clone()
        opt.step()
        for d, gr in zip(dev, grads):
            d['grad'].copy_(gr)
        ops.adam_step(dev, 0.9, 0.999, 1e-15, step)
    torch.cuda.synchronize()
    assert step[0].item() == 5
    for d, r in zip(dev, ref):
        np.testing.assert_allclose(d['param'].cpu().numpy(), r.detach().numpy(), rtol=2e-5, atol=2e-6)
        assert torch.all(d['grad'] == 0)
    np.testing.assert_array_equal(dev[0]['shadow_f16'].cpu().numpy(), dev[0]['param'].half().cpu().numpy())


def test_adam_grad_scaler_skip_on_inf():
    from bundlesdf_b200 import ops
    n = 1000
    p = torch.ones(n, device=DEV)
    seg = [dict(param=p, grad=torch.full((n,), 65536.0, device=DEV), exp_avg=torch.zeros(n, device=DEV), exp_avg_sq=torch.zeros(n, device=DEV), lr=0.1)]
    step = torch.zeros(8, dtype=torch.int32, device=DEV)      # [0] count, [1..7] library scratch (include/nof.h)
    scale = torch.tensor([65536.0, 0.0], device=DEV)
    inf = torch.ones(1, dtype=torch.int32, device=DEV)
    ops.adam_step(seg, 0.9, 0.999, 1e-15, step, scale, inf)
    torch.cuda.synchronize()
    assert torch.all(p == 1) and step[0].item() == 0 and scale[0].item() == 32768.0 and inf.item() == 0
    seg[0]['grad'].fill_(32768.0)                      # scaled gradient of 1.0
    ops.adam_step(seg, 0.9, 0.999, 1e-15, step, scale, inf)
    torch.cuda.synchronize()
    np.testing.assert_allclose(p.cpu().numpy(), 1 - 0.1, rtol=1e-6)
    assert step[0].item() == 1 and scale[1].item() == 1.0
