"""Generate golden vectors by running the REFERENCE's own compiled CUDA extensions (oracle/_ref/*.so, built by
oracle/build_ref.py from /root/reference/mycuda sources) on a B200, plus the reference's own Python that needs them.
Run on the GPU box only:

    gpurun -- 'python tests/golden/make_golden_gpu.py'      ->  gpurun_out/golden/ref_gpu_*.npz

then copy the files into tests/golden/ and commit them. Large inputs (hash tables) are not stored: they are regenerated
from the recorded torch CPU seeds by the tests (same torch build in the image => same stream).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)
OUT = os.path.join(REPO, 'gpurun_out', 'golden')


def load_ext(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def table(L, finest, log2T, seed, scale=0.5):
    from oracle import nof_oracle as O
    offsets, pls = O.grid_offsets(L, 16, finest, log2T)
    g = torch.Generator().manual_seed(seed)
    emb = (torch.rand(int(offsets[-1]), 2, generator=g) * 2 - 1) * scale
    return offsets, float(np.log2(pls)), float(pls), emb


def points(B, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, generator=g)
    x[0] = 0.0; x[1] = 1.0; x[2] = torch.tensor([0.5, 1.0, 0.0]); x[3] = torch.tensor([1.0001, 0.5, 0.5]); x[4] = torch.tensor([0.5, -1e-6, 0.5])
    return x


def main():
    os.makedirs(OUT, exist_ok=True)
    ge = load_ext('gridencoder_ref', os.path.join(REPO, 'oracle', '_ref', 'gridencoder', 'gridencoder_ref.so'))
    cm = load_ext('common_ref', os.path.join(REPO, 'oracle', '_ref', 'common', 'common_ref.so'))
    dev = 'cuda'
    sv = lambda name, **kw: np.savez_compressed(os.path.join(OUT, name), **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in kw.items()})

    # ---- 1. gridencoder forward/backward, fp32 and fp16, three level layouts (all-dense, dense+hash, mostly hash)
    for tag, L, finest, log2T in [('L4', 4, 128, 22), ('L16a', 16, 256, 19), ('L16b', 16, 512, 16)]:
        offsets, S, pls, emb = table(L, finest, log2T, 101)
        B = 1024
        x = points(B, 102)
        g = torch.Generator().manual_seed(103)
        gy = torch.randn(L, B, 2, generator=g) * 0.1
        for dt, dn in [(torch.float32, 'f32'), (torch.float16, 'f16')]:
            e = emb.to(dt).to(dev)
            out = torch.empty(L, B, 2, device=dev, dtype=dt)
            dy = torch.empty(B, L * 3 * 2, device=dev, dtype=dt)
            od = torch.from_numpy(offsets).to(dev)
            ge.grid_encode_forward(x.to(dev), e, od, out, B, 3, 2, L, S, 16, True, dy, 0, False)
            gemb = torch.zeros_like(e)
            gin = torch.zeros(B, 3, device=dev, dtype=dt)
            ge.grid_encode_backward(gy.to(dt).to(dev), x.to(dev), e, od, gemb, B, 3, 2, L, S, 16, True, dy, gin, 0, False)
            torch.cuda.synchronize()
            from bundlesdf_b200 import ops
            scales = ops.grid_level_scales(S, 16, L)
            nz = gemb.float().abs().sum(-1).nonzero().reshape(-1)
            sv(f'ref_gpu_grid_{tag}_{dn}.npz', L=L, finest=finest, log2T=log2T, table_seed=101, point_seed=102, grad_seed=103, B=B,
               out=out, dy_dx=dy, grad_inputs=gin, grad_emb_idx=nz, grad_emb_val=gemb[nz], scales=scales)

    # ---- 2. common.sampleRaysUniformOccupiedVoxels / postprocessOctreeRayTracing
    rng = np.random.default_rng(7)
    N, I, S = 257, 9, 96
    io = np.zeros((N, I, 2), np.float32)
    for r in range(N):
        k = rng.integers(0, I + 1)
        t = rng.uniform(0.5, 2.0)
        for j in range(k):
            a = t + rng.uniform(0.0, 0.2) * (rng.random() < 0.5)
            b = a + rng.uniform(0.01, 0.3)
            io[r, j] = (a, b); t = b
    total = (io[:, :, 1] - io[:, :, 0]).sum(-1, dtype=np.float32)
    zs = (rng.random((N, S), dtype=np.float32) * total[:, None] * 0.999).astype(np.float32)
    zv = torch.zeros(N, S, device=dev)
    cm.sampleRaysUniformOccupiedVoxels(torch.from_numpy(io).to(dev), torch.from_numpy(zs).to(dev), zv)
    n_rays = 80
    ray_index, dio = [], []
    for r in range(n_rays):
        if rng.random() < 0.2:
            continue
        k = rng.integers(1, 9); t = rng.uniform(0.5, 1.0)
        for j in range(k):
            a = t; b = a + (rng.uniform(0.0, 0.3) if rng.random() > 0.2 else 5e-5)
            if rng.random() < 0.05:
                a, b = b, a
            ray_index.append(r); dio.append((a, b)); t = b
    ray_index = np.array(ray_index, np.int64); dio = np.array(dio, np.float32)
    uniq, counts = np.unique(ray_index, return_counts=True)
    start = np.concatenate([[0], np.cumsum(counts[:-1])]).astype(np.int64)
    mi = int(counts.max())
    torch.cuda.set_device(0)
    pp = cm.postprocessOctreeRayTracing(torch.from_numpy(ray_index).to(dev), torch.from_numpy(dio).to(dev), torch.from_numpy(uniq).to(dev),
                                        torch.from_numpy(start).to(dev), mi, n_rays)
    torch.cuda.synchronize()
    sv('ref_gpu_common.npz', z_in_out=io, z_sampled=zs, z_vals=zv, ray_index=ray_index, depth_in_out=dio, unique_ids=uniq, start_poss=start,
       max_intersections=mi, n_rays=n_rays, padded=pp)

    # ---- 3. the reference's own Python on top of its extensions: GridEncoder under autocast, and
    #         NerfRunner.sample_rays_uniform_occupied_voxels / sample_rays_uniform on CUDA
    import ref_shims
    ref_py = '/root/reference' if os.path.isdir('/root/reference') else os.path.join(REPO, 'oracle', '_ref', 'py')
    nh, nr, U = ref_shims.import_reference(ref_py, mycuda_common=cm, mycuda_gridencoder=ge)
    sys.path.insert(0, os.path.join(ref_py, 'mycuda', 'torch_ngp_grid_encoder'))
    grid_mod = importlib.import_module('grid')
    torch.manual_seed(0)
    enc = grid_mod.GridEncoder(input_dim=3, n_levels=16, log2_hashmap_size=14, desired_resolution=256, base_resolution=16, level_dim=2).to(dev)
    with torch.no_grad():
        g = torch.Generator().manual_seed(201)
        enc.embeddings.copy_(((torch.rand(enc.embeddings.shape, generator=g) * 2 - 1) * 0.5).to(dev))
    xg = (points(512, 202) * 2 - 1).to(dev).requires_grad_(True)
    gyy = torch.randn(512, 32, generator=torch.Generator().manual_seed(203)).to(dev)
    res = {}
    for amp in (False, True):
        enc.zero_grad(); xg.grad = None
        with torch.cuda.amp.autocast(enabled=amp):
            y = enc(xg)
        (y.float() * gyy).sum().backward()
        res[f'y_amp{int(amp)}'] = y.float(); res[f'gx_amp{int(amp)}'] = xg.grad.clone(); res[f'gemb_amp{int(amp)}'] = enc.embeddings.grad.clone()
    sv('ref_gpu_gridencoder_module.npz', table_seed=201, point_seed=202, grad_seed=203, offsets=enc.offsets, per_level_scale=enc.per_level_scale, **res)

    cfg = dict(sc_factor=5.0, near=0.1, far=2.0, trunc=0.01, trunc_decay_type='', n_step=500, N_samples=64)
    fake = types.SimpleNamespace(cfg=cfg, global_step=0)
    fake.get_truncation = types.MethodType(nr.NerfRunner.get_truncation, fake)
    Nn = 129
    io2 = torch.from_numpy(io[:Nn]).to(dev)
    rays_d = torch.randn(Nn, 3, generator=torch.Generator().manual_seed(5)).to(dev)
    rays_d[:, 2] = -1.0
    depths = (torch.rand(Nn, generator=torch.Generator().manual_seed(6)) * 2.0 + 0.8).to(dev)
    depths[::9] = 99 * 5.0
    torch.manual_seed(77)
    z_vals, z_cont = nr.NerfRunner.sample_rays_uniform_occupied_voxels(fake, ray_ids=None, rays_d=rays_d, depths_in_out=io2, lindisp=False,
                                                                        perturb=True, depths=depths, N_samples=64)
    torch.manual_seed(77)
    t_rand = torch.rand(Nn, 64, device=dev)
    z_vals0, _ = nr.NerfRunner.sample_rays_uniform_occupied_voxels(fake, ray_ids=None, rays_d=rays_d, depths_in_out=io2, lindisp=False,
                                                                   perturb=False, depths=depths, N_samples=64)
    lin = torch.linspace(0., 1., steps=64, device=dev)
    lin192 = torch.linspace(0., 1., steps=192, device=dev)
    sv('ref_gpu_sample_occupied.npz', depths_in_out=io2, rays_d=rays_d, depths=depths, z_vals=z_vals, z_cont=z_cont, t_rand=t_rand,
       z_vals_noperturb=z_vals0, linspace64=lin, linspace192=lin192, sc_factor=5.0)
    print('golden GPU fixtures written to', OUT)


if __name__ == '__main__':
    main()
