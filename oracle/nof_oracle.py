"""CPU oracle for the Neural-Object-Field training step (TEST INFRASTRUCTURE — never the product path).

Plain PyTorch (fp32 by default, fp64 on request) restatement of the BundleSDF reference's hot path.
Every function cites the reference file:line it follows (paths relative to /root/reference).
Only tests/, bench.py's cpu_baseline / ``--impl reference`` leg and __graft_entry__.smoke() may
import this module; bundlesdf_b200/ must never import it (tests/test_no_oracle_in_product.py checks).

Parity pin status:
  * hash-grid encoder ............ pinned against the reference's own compiled gridencoder.cu run on a
                                   B200 (tests/golden/ref_gridencoder_*.npz, made by
                                   tests/golden/make_golden_gpu.py)
  * interval walk / nugget packing  pinned against the reference's compiled common.cu (same script)
  * SH, NeRFSmall, loss masks, stratified sampler, raw2outputs ... pinned against the reference's own
                                   Python (nerf_helpers.py / nerf_runner.py imported under shims on CPU,
                                   tests/golden/make_golden_cpu.py)
  * se3_exp_map (pytorch3d) and the octree ray trace (kaolin) are third-party, absent from
    /root/reference and from this image: restated from their published algorithms — PARITY UNPINNED
    for those two pieces (se3 is additionally checked against scipy.linalg.expm).
"""
import math
import numpy as np
import torch

# --------------------------------------------------------------------------------------------------
# Hash-grid encoder  (mycuda/torch_ngp_grid_encoder/grid.py:107-168, gridencoder.cu:47-246)
# --------------------------------------------------------------------------------------------------
PRIMES = (1, 2654435761, 805459861)   # gridencoder.cu:54


def grid_offsets(num_levels, base_res, finest_res, log2_hashmap_size, input_dim=3):
    """grid.py:110,125-138: per_level_scale, int32 offsets [L+1] (entries, not floats)."""
    per_level_scale = np.exp2(np.log2(finest_res / base_res) / (num_levels - 1))
    max_params = 2 ** log2_hashmap_size
    offsets, offset = [], 0
    for i in range(num_levels):
        resolution = int(np.ceil(base_res * per_level_scale ** i))
        params_in_level = min(max_params, (resolution + 1) ** input_dim)
        params_in_level = int(np.ceil(params_in_level / 8) * 8)
        offsets.append(offset)
        offset += params_in_level
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32), float(per_level_scale)


def level_scale_res(level, S, H):
    """gridencoder.cu:155-156 in fp32: scale = exp2f(level*S)*H - 1 (nvcc contracts to one FMA);
    resolution = ceil(scale)+1."""
    e = np.exp2(np.float32(np.float32(level) * np.float32(S)), dtype=np.float32)
    scale = np.float32(np.float64(e) * np.float64(H) - 1.0)     # fma: single rounding
    res = int(np.ceil(scale)) + 1
    return scale, res


def _grid_index(cx, cy, cz, hashmap_size, resolution):
    """gridencoder.cu:66-83 (align_corners=False, gridtype hash). c* are int64 tensors holding uint32 values."""
    stride = 1
    index = torch.zeros_like(cx)
    for c in (cx, cy, cz):
        if stride <= hashmap_size:
            index = index + c * stride
            stride *= (resolution + 1)
    if stride > hashmap_size:
        M = 0xFFFFFFFF
        index = ((cx * PRIMES[0]) & M) ^ ((cy * PRIMES[1]) & M) ^ ((cz * PRIMES[2]) & M)
    else:
        index = index & 0xFFFFFFFF
    return index % hashmap_size


def grid_encode(x01, embeddings, offsets, S, H, exact_fma=True, want_dydx=False, scales=None):
    """x01 [B,3] in [0,1] (fp32), embeddings [sO,C], offsets int array [L+1].
    Returns out [B, L*C] (and dy_dx [B,L,3,C] = d out / d x01, gridencoder.cu:202-245).
    Differentiable w.r.t. embeddings and (through the interpolation weights) x01.
    exact_fma emulates the FMA contractions nvcc applies to the reference kernel so the fp32 forward is
    bit-comparable with the compiled reference. scales: optional per-level `scale` values as evaluated on the
    device (CUDA's exp2f differs from libm by an ulp at some levels; tests read them back from the GPU or from
    the golden fixtures) — with them the fp32 forward is bit-identical to the compiled reference kernel."""
    B = x01.shape[0]
    L = len(offsets) - 1
    C = embeddings.shape[1]
    dt = embeddings.dtype
    oob = ((x01 < 0) | (x01 > 1)).any(dim=-1)                      # gridencoder.cu:128-152
    outs, dydxs = [], []
    for l in range(L):
        hsize = int(offsets[l + 1] - offsets[l])
        scale, res = level_scale_res(l, S, H)
        if scales is not None:
            scale = np.float32(scales[l])
            res = int(np.ceil(scale)) + 1
        if exact_fma and x01.dtype == torch.float32:
            pos = (x01.double() * float(scale) + 0.5).float()      # fmaf(x, scale, 0.5)
        else:
            pos = x01 * float(scale) + 0.5
        pg = torch.floor(pos.detach())
        frac = pos - pg                                             # exact in fp32
        pg = pg.long().clamp(min=0)                                 # uint32 cast; oob rows are masked below
        tab = embeddings[int(offsets[l]):int(offsets[l + 1])]
        acc = torch.zeros(B, C, dtype=dt, device=x01.device)
        feats = []
        for idx in range(8):
            w = torch.ones(B, dtype=x01.dtype, device=x01.device)
            cs = []
            for d in range(3):
                if (idx >> d) & 1:
                    w = w * frac[:, d]
                    cs.append(pg[:, d] + 1)
                else:
                    w = w * (1 - frac[:, d])
                    cs.append(pg[:, d])
            gi = _grid_index(cs[0], cs[1], cs[2], hsize, res)
            f = tab[gi]
            feats.append(f)
            if exact_fma and dt == torch.float32:
                acc = (w.double()[:, None] * f.double() + acc.double()).float()   # fmaf(w, g, acc)
            else:
                acc = acc + w[:, None].to(dt) * f
        acc = torch.where(oob[:, None], torch.zeros_like(acc), acc)
        outs.append(acc)
        if want_dydx:
            dl = []
            for gd in range(3):
                rg = torch.zeros(B, C, dtype=dt, device=x01.device)
                others = [d for d in range(3) if d != gd]
                for idx in range(4):
                    w = torch.full((B,), float(scale), dtype=x01.dtype, device=x01.device)
                    base = 0
                    for nd, d in enumerate(others):
                        if (idx >> nd) & 1:
                            w = w * frac[:, d]
                            base |= (1 << d)
                        else:
                            w = w * (1 - frac[:, d])
                    left, right = feats[base], feats[base | (1 << gd)]
                    rg = rg + w[:, None].to(dt) * (right - left)
                dl.append(torch.where(oob[:, None], torch.zeros_like(rg), rg))
            dydxs.append(torch.stack(dl, dim=1))                     # [B,3,C]
    out = torch.cat(outs, dim=-1)                                    # [B, L*C] (grid.py:64)
    if want_dydx:
        return out, torch.stack(dydxs, dim=1)                        # [B,L,3,C]
    return out


# --------------------------------------------------------------------------------------------------
# Pose correction (nerf_helpers.py:127-154; pytorch3d.transforms.se3_exp_map — third-party, restated)
# --------------------------------------------------------------------------------------------------
def _hat(v):
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    o = torch.zeros_like(x)
    return torch.stack([torch.stack([o, -z, y], -1), torch.stack([z, o, -x], -1), torch.stack([-y, x, o], -1)], 1)


def se3_exp_map(log_transform, eps=1e-4):
    """pytorch3d se3_exp_map semantics: input [F,6]=(v|omega); returns the TRANSPOSED 4x4 (row-vector
    convention), exactly like pytorch3d, so PoseArray's .permute(0,2,1) (nerf_helpers.py:150) applies."""
    v, w = log_transform[:, :3], log_transform[:, 3:]
    nrms = (w * w).sum(1)
    ang = torch.clamp(nrms, eps).sqrt()
    inv = 1.0 / ang
    fac1 = inv * ang.sin()
    fac2 = inv * inv * (1.0 - ang.cos())
    K = _hat(w)
    K2 = torch.bmm(K, K)
    I = torch.eye(3, dtype=w.dtype, device=w.device)[None]
    R = fac1[:, None, None] * K + fac2[:, None, None] * K2 + I
    V = I + K * ((1 - torch.cos(ang)) / (ang ** 2))[:, None, None] + K2 * ((ang - torch.sin(ang)) / (ang ** 3))[:, None, None]
    T = torch.bmm(V, v[:, :, None])[:, :, 0]
    out = torch.zeros(len(v), 4, 4, dtype=w.dtype, device=w.device)
    out[:, :3, :3] = R
    out[:, :3, 3] = T
    out[:, 3, 3] = 1.0
    return out.permute(0, 2, 1)


def pose_matrices(pose_data, max_trans, max_rot_deg):
    """nerf_helpers.py:143-154 for ids = arange(F): frame 0 forced to identity. Returns [F,4,4]."""
    theta = torch.tanh(pose_data)
    trans = theta[:, :3] * max_trans
    rot = theta[:, 3:6] * max_rot_deg / 180.0 * np.pi
    Ts = se3_exp_map(torch.cat((trans, rot), dim=-1)).permute(0, 2, 1)
    eye = torch.eye(4, dtype=Ts.dtype, device=Ts.device)[None]
    mask = torch.ones(len(Ts), 1, 1, dtype=torch.bool, device=Ts.device)
    mask[0] = False
    return torch.where(mask, Ts, eye.expand_as(Ts))


# --------------------------------------------------------------------------------------------------
# SH view encoding (nerf_helpers.py:67-105, degree 3) and NeRFSmall (nerf_helpers.py:243-321)
# --------------------------------------------------------------------------------------------------
SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)


def sh_encode_deg3(d):
    x, y, z = d.unbind(-1)
    xx, yy, zz = x * x, y * y, z * z
    xy, yz, xz = x * y, y * z, x * z
    return torch.stack([torch.full_like(x, SH_C0), -SH_C1 * y, SH_C1 * z, -SH_C1 * x,
                        SH_C2[0] * xy, SH_C2[1] * yz, SH_C2[2] * (2.0 * zz - xx - yy), SH_C2[3] * xz,
                        SH_C2[4] * (xx - yy)], dim=-1)


def init_mlp(enc_dim, view_dim, seed=0, dtype=torch.float32):
    """Weights with nn.Linear default init in the module order of NeRFSmall (nerf_helpers.py:255-294),
    sigma_net last bias = 0.1 (:272). Returns dict with the reference's state_dict key names."""
    g = torch.Generator().manual_seed(seed)
    def lin(o, i):
        bound = 1.0 / math.sqrt(i)
        W = (torch.rand(o, i, generator=g, dtype=dtype) * 2 - 1) * bound     # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(i), 1/sqrt(i))
        b = (torch.rand(o, generator=g, dtype=dtype) * 2 - 1) * bound
        return W, b
    p = {}
    p['sigma_net.0.weight'], p['sigma_net.0.bias'] = lin(64, enc_dim)
    p['sigma_net.2.weight'], p['sigma_net.2.bias'] = lin(16, 64)
    p['sigma_net.2.bias'] = torch.full((16,), 0.1, dtype=dtype)
    p['color_net.0.weight'], p['color_net.0.bias'] = lin(64, view_dim + 15)
    p['color_net.2.weight'], p['color_net.2.bias'] = lin(64, 64)
    p['color_net.4.weight'], p['color_net.4.bias'] = lin(3, 64)
    return p


def _q(t, half):
    """Emulate an fp16 tensor-core operand: round to fp16, compute in fp32 (accumulate fp32)."""
    return t.half().float() if half else t


def mlp_forward(p, enc, views, half=False):
    """NeRFSmall.forward (nerf_helpers.py:305-321). enc [P,E]; views [P, ff+9]. Returns [P,4] = (rgb logits, sdf).
    half=True emulates autocast (fp16 operands, fp32 accumulate, fp16 layer outputs)."""
    F = torch.nn.functional
    def lin(x, W, b):
        y = F.linear(_q(x, half), _q(W, half), _q(b, half))
        return _q(y, half)
    h = torch.relu(lin(enc, p['sigma_net.0.weight'], p['sigma_net.0.bias']))
    h = lin(h, p['sigma_net.2.weight'], p['sigma_net.2.bias'])
    sdf, geo = h[..., 0], h[..., 1:]
    c = torch.cat([views, geo], dim=-1)
    c = torch.relu(lin(c, p['color_net.0.weight'], p['color_net.0.bias']))
    c = torch.relu(lin(c, p['color_net.2.weight'], p['color_net.2.bias']))
    c = lin(c, p['color_net.4.weight'], p['color_net.4.bias'])
    return torch.cat([c, sdf[..., None]], dim=-1)


def mlp_forward_sdf(p, enc, half=False):
    """NeRFSmall.forward_sdf (nerf_helpers.py:296-302)."""
    F = torch.nn.functional
    h = torch.relu(_q(F.linear(_q(enc, half), _q(p['sigma_net.0.weight'], half), _q(p['sigma_net.0.bias'], half)), half))
    h = _q(F.linear(_q(h, half), _q(p['sigma_net.2.weight'], half), _q(p['sigma_net.2.bias'], half)), half)
    return h[..., 0]


# --------------------------------------------------------------------------------------------------
# Occupancy (nerf_runner.py:436-476) and ray/voxel intervals (Utils.py:443-475 + kaolin raytrace,
# third-party & absent: restated as an exact voxel DDA — PARITY UNPINNED; common.cu:129-149)
# --------------------------------------------------------------------------------------------------
def octree_levels(cfg):
    """nerf_runner.py:444-447 (max_level) and :1058-1059 (ray-tracing level)."""
    sc = cfg['sc_factor']
    max_level = int(np.ceil(np.log2(2.0 / (cfg['octree_smallest_voxel_size'] * sc))))
    level = int(np.floor(np.log2(2.0 / (cfg['octree_raytracing_voxel_size'] * sc))))
    return max_level, level


def build_occupancy(pts, cfg):
    """nerf_runner.py:443-476: quantise cloud at max_level, dilate by the 27-neighbourhood
    dilate_radius times, clip centres to [-1,1], re-quantise (kaolin quantize_points, Utils.py:362) and
    coarsen to the ray-tracing level. pts np [M,3] in [-1,1]. Returns (occ bool np [n,n,n] indexed [x,y,z], level)."""
    max_level, level = octree_levels(cfg)
    vox = 2.0 / (2 ** max_level)
    dilate_radius = max(1, int(np.ceil(cfg['octree_dilate_size'] / cfg['octree_smallest_voxel_size'])))
    coords = np.floor((np.asarray(pts, np.float32) + 1) / np.float32(vox)).astype(np.int64)
    coords = np.unique(coords, axis=0)
    shifts = np.array([[dx, dy, dz] for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1)], np.int64)
    for _ in range(dilate_radius):
        coords = np.unique((coords[None] + shifts[:, None]).reshape(-1, 3), axis=0)
    centers = np.clip((coords + 0.5) * vox - 1, -1, 1)
    n_max = 2 ** max_level
    q = np.clip(np.floor((centers + 1) / 2 * n_max), 0, n_max - 1).astype(np.int64)   # kaolin quantize_points
    shift = max_level - level
    if shift >= 0:
        q = q >> shift
    else:   # tracing level finer than the octree: cannot happen with shipped cfg (same voxel size)
        raise ValueError('ray tracing level deeper than octree max_level')
    n = 2 ** level
    occ = np.zeros((n, n, n), dtype=bool)
    occ[q[:, 0], q[:, 1], q[:, 2]] = True
    return occ, level


def ray_trace_intervals(occ, rays_o, rays_d, i_max=None):
    """Per occupied voxel pierced, front to back, (t_in,t_out) of Euclidean travel along the UNIT dir
    (kaolin unbatched_raytrace(return_depth, with_exit) semantics, Utils.py:457), then the reference's
    packing rule common.cu:137-148: stop at an entry whose t_in==0 or t_out==0, skip t_in>t_out and
    |t_out-t_in|<1e-4, pad with zeros. fp32 arithmetic, one rounding per operation (the CUDA sampler is
    compiled without FMA contraction for this code) so results are bit-comparable.
    occ np bool [n,n,n]; rays_o, rays_d np fp32 [N,3]. Returns np fp32 [N,I,2] (I = i_max or max count, >=1)."""
    f32 = np.float32
    n = occ.shape[0]
    cell = f32(2.0) / f32(n)
    N = len(rays_o)
    out = []
    for r in range(N):
        o = rays_o[r].astype(f32)
        d = rays_d[r].astype(f32)
        lst = []
        with np.errstate(divide='ignore', invalid='ignore'):
            inv = f32(1.0) / d
        # slab test against [-1,1]^3
        t0, t1 = f32(0.0), f32(np.inf)
        hit = True
        for a in range(3):
            if d[a] == 0:
                if o[a] < -1 or o[a] > 1:
                    hit = False
                continue
            ta = (f32(-1.0) - o[a]) * inv[a]
            tb = (f32(1.0) - o[a]) * inv[a]
            lo, hi = (ta, tb) if ta <= tb else (tb, ta)
            t0 = max(t0, lo)
            t1 = min(t1, hi)
        if hit and t0 < t1:
            # start cell from the entry point (nudged to the middle of the first cell crossing)
            ix = [0, 0, 0]
            step = [0, 0, 0]
            for a in range(3):
                p = o[a] + t0 * d[a]
                c = int(np.floor((p + f32(1.0)) / cell))
                ix[a] = min(max(c, 0), n - 1)
                step[a] = 1 if d[a] > 0 else (-1 if d[a] < 0 else 0)
            t_in = t0
            guard = 0
            while guard < 3 * n + 3:
                guard += 1
                # exit time of the current cell: nearest of the three exit planes
                t_out = f32(np.inf)
                ax = -1
                for a in range(3):
                    if step[a] == 0:
                        continue
                    plane = f32(ix[a] + (1 if step[a] > 0 else 0)) * cell - f32(1.0)
                    ta = (plane - o[a]) * inv[a]
                    if ta < t_out:
                        t_out = ta
                        ax = a
                if ax < 0:
                    break
                t_out = min(t_out, t1)
                if occ[ix[0], ix[1], ix[2]]:
                    lst.append((t_in, t_out))
                ix[ax] += step[ax]
                if ix[ax] < 0 or ix[ax] >= n or t_out >= t1:
                    break
                t_in = t_out
        # common.cu:137-148 packing
        packed = []
        for (a, b) in lst:
            if a == 0 or b == 0:
                break
            if a > b:
                continue
            if float(abs(f32(b - a))) < 1e-4:          # float diff promoted to double vs the literal 1e-4 (common.cu:140)
                continue
            packed.append((a, b))
        out.append(packed)
    I = max(1, max(len(p) for p in out)) if i_max is None else i_max
    res = np.zeros((N, I, 2), dtype=f32)
    for r, p in enumerate(out):
        for k, (a, b) in enumerate(p[:I]):
            res[r, k, 0] = a
            res[r, k, 1] = b
    return res


def ray_trace_intervals_merge(occ, rays_o, rays_d, i_max=None):
    """Same result as ray_trace_intervals, bit for bit, WITHOUT the sequential walk: the design check for a warp-parallel
    ray_march kernel. The exit time of a cell through the k-th plane of axis a, T_a[k] = (plane(ix0_a + k step_a) -
    o_a) * inv_a, is a closed form of k (the walk never accumulates t), non-decreasing in k (every operation is a monotone rounding),
    so the walk is the merge of three sorted lists ordered by (T, axis) — `ta < t_out` with axes tried in order 0,1,2 is exactly that
    tie rule. Step m leaves through event m; its cell is the start cell advanced by the per-axis event counts before m; t_in(m) =
    min(T[m-1], t1) (t0 for m = 0); the walk ends at the first event that leaves the grid (the last crossing of its axis) or reaches
    t1. All of it is data-parallel over events; only the packing rule (a stop flag and a running count) needs a prefix scan."""
    f32 = np.float32
    n = occ.shape[0]
    cell = f32(2.0) / f32(n)
    N = len(rays_o)
    out = []
    for r in range(N):
        o = rays_o[r].astype(f32)
        d = rays_d[r].astype(f32)
        with np.errstate(divide='ignore', invalid='ignore'):
            inv = f32(1.0) / d
        t0, t1 = f32(0.0), f32(np.inf)
        hit = True
        for a in range(3):
            if d[a] == 0:
                if o[a] < -1 or o[a] > 1:
                    hit = False
                continue
            ta = (f32(-1.0) - o[a]) * inv[a]
            tb = (f32(1.0) - o[a]) * inv[a]
            lo, hi = (ta, tb) if ta <= tb else (tb, ta)
            t0 = max(t0, lo)
            t1 = min(t1, hi)
        packed = []
        if hit and t0 < t1:
            ix0 = np.zeros(3, np.int64)
            step = np.zeros(3, np.int64)
            for a in range(3):
                p = o[a] + t0 * d[a]
                c = int(np.floor((p + f32(1.0)) / cell))
                ix0[a] = min(max(c, 0), n - 1)
                step[a] = 1 if d[a] > 0 else (-1 if d[a] < 0 else 0)
            Ts, axs, ks, last = [], [], [], []
            for a in range(3):
                if step[a] == 0:
                    continue
                K = int(n - ix0[a]) if step[a] > 0 else int(ix0[a] + 1)          # crossings until the walk leaves the grid on this axis
                k = np.arange(K, dtype=np.int64)
                ixa = ix0[a] + k * step[a]
                plane = (ixa + (1 if step[a] > 0 else 0)).astype(f32) * cell - f32(1.0)
                T = ((plane - o[a]).astype(f32) * inv[a]).astype(f32)
                Ts.append(T); axs.append(np.full(K, a)); ks.append(k); last.append(k == K - 1)
            if Ts:
                T = np.concatenate(Ts); ax = np.concatenate(axs); kk = np.concatenate(ks); is_last = np.concatenate(last)
                order = np.lexsort((kk, ax, T))                                  # by T, ties: lower axis first, then k
                T, ax, is_last = T[order], ax[order], is_last[order]
                M = len(T)
                t_out = np.minimum(T, t1)
                t_in = np.concatenate([[t0], t_out[:-1]]).astype(f32)
                cnt = np.zeros((M, 3), np.int64)                                 # events of each axis before step m
                for a in range(3):
                    cnt[:, a] = np.concatenate([[0], np.cumsum(ax == a)[:-1]])
                cells = ix0[None, :] + cnt * step[None, :]
                stop = is_last | (t_out >= t1)
                m_end = int(np.argmax(stop)) if stop.any() else M - 1
                m_end = min(m_end, 3 * n + 2)                                    # the walk's guard (never binding: M <= 3n)
                sel = np.arange(M) <= m_end
                o_m = occ[cells[:, 0].clip(0, n - 1), cells[:, 1].clip(0, n - 1), cells[:, 2].clip(0, n - 1)] & sel
                # packing rule (common.cu:137-148) over the occupied steps, in order
                zero = o_m & ((t_in == 0) | (t_out == 0))
                first_zero = int(np.argmax(zero)) if zero.any() else M
                keep = o_m & (np.arange(M) < first_zero) & ~(t_in > t_out) & ~(np.abs((t_out - t_in).astype(f32)).astype(np.float64) < 1e-4)
                packed = [(a_, b_) for a_, b_ in zip(t_in[keep], t_out[keep])]
        out.append(packed)
    I = max(1, max(len(p) for p in out)) if i_max is None else i_max
    res = np.zeros((N, I, 2), dtype=f32)
    for r, p in enumerate(out):
        for k, (a, b) in enumerate(p[:I]):
            res[r, k, 0] = a
            res[r, k, 1] = b
    return res


def postprocess_octree_ray_tracing(ray_index, depth_in_out, unique_ids, start_poss, max_intersections, n_rays):
    """common.cu:129-167 verbatim semantics on CPU (numpy)."""
    out = np.zeros((n_rays, max_intersections, 2), np.float32)
    for u in range(len(unique_ids)):
        i_ray = int(unique_ids[u])
        k = 0
        for i in range(int(start_poss[u]), len(ray_index)):
            if ray_index[i] != i_ray:
                break
            a, b = depth_in_out[i]
            if a == 0 or b == 0:
                break
            if a > b:
                continue
            if float(abs(np.float32(b - a))) < 1e-4:
                continue
            out[i_ray, k] = (a, b)
            k += 1
    return out


# --------------------------------------------------------------------------------------------------
# Sampling (nerf_runner.py:67-87, 979-1011, 1063-1081; common.cu:41-105)
# --------------------------------------------------------------------------------------------------
def sample_rays_uniform(n_samples, near, far, t_rand=None):
    """nerf_runner.py:67-87 (lindisp False). near/far [N,1]; t_rand [N,S] injected uniform randoms
    (None -> perturb off)."""
    t_vals = torch.linspace(0., 1., steps=n_samples, dtype=near.dtype).reshape(1, -1)
    z_vals = near * (1. - t_vals) + far * t_vals
    if t_rand is not None:
        mids = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
        upper = torch.cat([mids, z_vals[..., -1:]], -1)
        lower = torch.cat([z_vals[..., :1], mids], -1)
        z_vals = lower + (upper - lower) * t_rand
        z_vals = torch.minimum(torch.maximum(z_vals, near), far)          # torch.clip(z, near, far)
    return z_vals


def interval_walk(z_in_out, z_sampled):
    """common.cu:41-105 on CPU. z_in_out np [N,I,2], z_sampled np [N,S] -> (z_vals np [N,S], err flag).
    Where the reference prints and spins forever (:66-71,:87-92) we clamp to the last interval's exit and
    set err (the CUDA path does the same and raises a device error flag)."""
    f32 = np.float32
    N, S = z_sampled.shape
    I = z_in_out.shape[1]
    z_vals = np.zeros((N, S), f32)
    err = False
    eps = f32(1e-4)
    for r in range(N):
        io = z_in_out[r]
        if io[0, 0] == 0:
            continue
        for s in range(S):
            rem = f32(z_sampled[r, s])
            k = 0
            while True:
                if k >= I:
                    if not rem <= eps:
                        err = True
                    z_vals[r, s] = io[I - 1, 1]
                    break
                if io[k, 0] == 0:
                    if not (rem <= eps and k >= 1):
                        err = True
                    z_vals[r, s] = io[k - 1, 1] if k >= 1 else 0
                    break
                blen = f32(io[k, 1] - io[k, 0])
                if rem <= blen:
                    z_vals[r, s] = f32(io[k, 0] + rem)
                    break
                rem = f32(rem - blen)
                k += 1
    return z_vals, err


def linspace01_cuda(S):
    """torch.linspace(0,1,S) as evaluated by torch's CUDA kernel (what the reference actually runs,
    nerf_runner.py:74): step=(end-start)/(steps-1); idx<steps/2: start+step*idx, else end-step*(steps-idx-1),
    each contracted to a single FMA. (The CPU kernel, used by the golden fixture, differs in the last ulp.)"""
    f32 = np.float32
    if S == 1:
        return np.zeros(1, f32)
    step = f32(1.0) / f32(S - 1)
    i = np.arange(S)
    lo = (step * i.astype(f32)).astype(f32)
    hi = (1.0 - np.float64(step) * (S - i - 1).astype(np.float64)).astype(f32)
    return np.where(i < S // 2, lo, hi).astype(f32)


def stratified_np(S, near, far, t_rand):
    """sample_rays_uniform (nerf_runner.py:67-87) in numpy fp32, one rounding per operation, CUDA linspace."""
    f32 = np.float32
    t = linspace01_cuda(S)[None, :]
    near = near.astype(f32).reshape(-1, 1)
    far = far.astype(f32).reshape(-1, 1)
    z = (near * (f32(1.0) - t)).astype(f32) + (far * t).astype(f32)
    z = z.astype(f32)
    if t_rand is not None:
        mids = (f32(0.5) * (z[:, 1:] + z[:, :-1]).astype(f32)).astype(f32)
        upper = np.concatenate([mids, z[:, -1:]], -1)
        lower = np.concatenate([z[:, :1], mids], -1)
        z = (lower + ((upper - lower).astype(f32) * t_rand.astype(f32)).astype(f32)).astype(f32)
        z = np.minimum(np.maximum(z, near), far)
    return z.astype(f32)


def rays_world_np(batch, tf12):
    """Unit camera dir, world origin, world unit dir per ray (nerf_runner.py:1045-1057) in the exact fp32 operation
    order of the CUDA sampler: nrm=sqrt((dx*dx+dy*dy)+dz*dz); u=d/nrm; dw_i=(R_i0*u0+R_i1*u1)+R_i2*u2.
    batch np [N,>=9], tf12 np [F,12]."""
    f32 = np.float32
    d = batch[:, 0:3].astype(f32)
    nrm = np.sqrt(((d[:, 0] * d[:, 0]).astype(f32) + (d[:, 1] * d[:, 1]).astype(f32)).astype(f32) + (d[:, 2] * d[:, 2]).astype(f32)).astype(f32)
    u = (d / nrm[:, None]).astype(f32)
    T = tf12[batch[:, 8].astype(np.int64)].astype(f32).reshape(-1, 3, 4)
    o = T[:, :, 3].copy()
    dw = np.stack([(((T[:, a, 0] * u[:, 0]).astype(f32) + (T[:, a, 1] * u[:, 1]).astype(f32)).astype(f32)
                    + (T[:, a, 2] * u[:, 2]).astype(f32)).astype(f32) for a in range(3)], -1)
    return u, o, dw


def sample_along_rays(depths_in_out, unit_dirs_cam, depth, cfg, trunc, t_rand):
    """nerf_runner.py:979-1011 (occupied-voxel sampling) + :1063-1081 (around-depth samples) in numpy fp32 with the
    operation order of the CUDA sampler. depths_in_out np [N,I,2] travel times; unit_dirs_cam np [N,3];
    depth np [N]; t_rand np [N, S_occ+S_depth] or None (perturb off). Returns (z_vals np [N,S], err)."""
    f32 = np.float32
    N, I = depths_in_out.shape[:2]
    sc = cfg['sc_factor']
    S_occ, S_d = cfg['N_samples'], cfg['N_samples_around_depth']
    absz = np.abs(unit_dirs_cam[:, 2]).astype(f32)
    z_io = (depths_in_out.astype(f32) * absz[:, None, None]).astype(f32)                 # :990
    depth = depth.astype(f32)
    near_sc, far_sc = f32(cfg['near'] * sc), f32(cfg['far'] * sc)
    valid_depth = (depth >= near_sc) & (depth <= far_sc)
    zmax = (depth + f32(trunc)).astype(f32)
    clip_ok = valid_depth[:, None] & (z_io > 0).all(-1)                                   # :994-995
    clipped = np.minimum(np.maximum(z_io, f32(0)), zmax[:, None, None])
    z_clip = np.where(clip_ok[..., None], clipped, z_io).astype(f32)

    def seq_total(io):
        tot = np.zeros(N, f32)
        for k in range(I):
            tot = (tot + (io[:, k, 1] - io[:, k, 0]).astype(f32)).astype(f32)
        return tot

    tr_occ = None if t_rand is None else t_rand[:, :S_occ]
    z_cont = stratified_np(S_occ, np.zeros(N, f32), seq_total(z_clip), tr_occ)
    z_occ, err = interval_walk(z_clip, z_cont)
    if S_d > 0:
        tr_d = None if t_rand is None else t_rand[:, S_occ:]
        nd = (depth - f32(trunc)).astype(f32)
        fd = (depth + (f32(trunc) * f32(cfg['neg_trunc_ratio'])).astype(f32)).astype(f32)
        z_ad = stratified_np(S_d, nd, fd, tr_d)
        if (~valid_depth).any():                                                          # :1074-1076
            z_c2 = stratified_np(S_d, np.zeros(N, f32), seq_total(z_io), tr_d)
            z_inv, e2 = interval_walk(z_io, z_c2)
            z_ad = np.where(valid_depth[:, None], z_ad, z_inv)
            err = err or e2
        z_occ = np.concatenate([z_occ, z_ad.astype(f32)], -1)
    return z_occ.astype(f32), err


# --------------------------------------------------------------------------------------------------
# Compositing + losses (nerf_runner.py:1132-1169, 679-752; nerf_helpers.py:367-399)
# --------------------------------------------------------------------------------------------------
def composite_weights(z_vals, depth, trunc, cfg):
    """sdf2weights, nerf_runner.py:1152-1161 (does NOT depend on the predicted sdf)."""
    sc = cfg['sc_factor']
    d = depth.view(-1, 1)
    s = (d - z_vals) / trunc
    w = torch.sigmoid(s * cfg['sdf_lambda']) * torch.sigmoid(-s * cfg['sdf_lambda'])
    invalid = (d > cfg['far'] * sc).reshape(-1)
    mask = (z_vals - d <= trunc * cfg['neg_trunc_ratio']) & (z_vals - d >= -trunc)
    w = torch.where(invalid[:, None], torch.zeros_like(w), w * mask)
    return w / (w.sum(dim=-1, keepdim=True) + 1e-10)


def step_losses(raw, z_vals, valid_samples, batch, trunc, cfg, pose_data=None, feature_data=None):
    """train_loop loss assembly, nerf_runner.py:679-752 with raw2outputs (:1163-1167). batch [N,12] rows:
    dir(3) rgb(3) depth mask frame_id type near far (make_frame_rays :259-300)."""
    sc = cfg['sc_factor']
    N, S = z_vals.shape
    target_s = batch[:, 3:6]
    target_d = batch[:, 6]
    frame_ids = batch[:, 8]
    ray_type = batch[:, 9]
    sdf = raw[..., 3]
    w = composite_weights(z_vals, target_d, trunc, cfg)
    w = torch.where(valid_samples, w, torch.zeros_like(w))
    rgb = (w[..., None] * torch.sigmoid(raw[..., :3])).sum(dim=-2)
    valid_rays = valid_samples.any(dim=-1) & (ray_type == 0)
    ray_w = torch.where(frame_ids == 0, torch.full_like(target_d, float(cfg['first_frame_weight'])), torch.ones_like(target_d))
    ray_w = ray_w * valid_rays
    sample_w = ray_w.view(N, 1).expand(-1, S) * valid_samples
    rgb_loss = cfg['rgb_weight'] * ((rgb - target_s) ** 2 * ray_w.view(-1, 1)).mean()
    sample_w = torch.where((ray_type == 1)[:, None], torch.zeros_like(sample_w), sample_w)
    td = target_d.reshape(-1, 1).expand(-1, S)
    valid_depth = (td >= cfg['near'] * sc) & (td <= cfg['far'] * sc)
    front = z_vals < td - trunc
    back = z_vals > td + trunc * cfg['neg_trunc_ratio']
    sdf_mask = (~front) & (~back) & valid_depth
    m_fs = (td > cfg['far'] * sc) & (sdf < cfg['fs_sdf'])
    fs_loss = torch.mean(((sdf - cfg['fs_sdf']) * m_fs) ** 2 * sample_w) * 0.5
    m_e = front & (td <= cfg['far'] * sc) & (sdf < 1)
    fs_loss = fs_loss + torch.mean(torch.abs(sdf - 1) * m_e * sample_w) * cfg['empty_weight']
    sdf_loss = torch.mean(((z_vals + sdf * trunc) * sdf_mask - td * sdf_mask) ** 2 * sample_w) * 0.5
    fs_loss = fs_loss * cfg['fs_weight']
    sdf_loss = sdf_loss * cfg['trunc_weight']
    loss = rgb_loss + fs_loss + sdf_loss
    out = {'rgb_loss': rgb_loss, 'fs_loss': fs_loss, 'sdf_loss': sdf_loss, 'rgb_map': rgb, 'weights': w}
    if cfg.get('fs_rgb_weight', 0) > 0:
        fs_rgb = ((((torch.sigmoid(raw[..., :3]) - 1) * front[..., None]) ** 2) * sample_w[..., None]).mean()
        loss = loss + fs_rgb * cfg['fs_rgb_weight']
        out['fs_rgb_loss'] = fs_rgb
    if feature_data is not None:
        reg = cfg['feature_reg_weight'] * (feature_data ** 2).mean()
        loss = loss + reg
        out['reg_features'] = reg
    if pose_data is not None and cfg.get('pose_reg_weight', 0) > 0:
        pr = cfg['pose_reg_weight'] * pose_data[1:].norm()
        loss = loss + pr
        out['pose_reg'] = pr
    out['loss'] = loss
    return out


# --------------------------------------------------------------------------------------------------
# One training step (nerf_runner.py:1014-1088 render_rays, :1227-1304 run_network, :679-763 train_loop)
# --------------------------------------------------------------------------------------------------
def get_truncation(cfg, global_step=0):
    """nerf_runner.py:663-676."""
    if cfg.get('trunc_decay_type', '') == 'linear':
        t = cfg['trunc_start'] - (cfg['trunc_start'] - cfg['trunc']) * float(global_step) / cfg['n_step']
    elif cfg.get('trunc_decay_type', '') == 'exp':
        lamb = np.log(cfg['trunc'] / cfg['trunc_start']) / (cfg['n_step'] / 4)
        t = max(cfg['trunc_start'] * np.exp(global_step * lamb), cfg['trunc'])
    else:
        t = cfg['trunc']
    return t * cfg['sc_factor']


def frame_transforms(params, c2w, cfg):
    """tf = pose_array.get_matrices(ids) @ c2w[ids]  (nerf_runner.py:1051-1053), for all frames: [F,4,4]."""
    if params.get('pose_data') is not None:
        dT = pose_matrices(params['pose_data'], cfg['max_trans'] * cfg['sc_factor'], cfg['max_rot'])
        return dT @ c2w
    return c2w


def forward_step(params, batch, c2w, occ, cfg, t_rand_occ=None, t_rand_depth=None, global_step=0, half=False,
                 z_vals=None):
    """Full forward of one train step on CPU. params: 'embeddings', MLP keys, optional 'pose_data' [F,6],
    'feature_data' [F,ff]; 'offsets' (np int32), 'S' (log2 per-level scale), 'H'. Returns dict with loss etc."""
    N = batch.shape[0]
    trunc = get_truncation(cfg, global_step)
    rays_d = batch[:, 0:3]
    viewdirs = rays_d / rays_d.norm(dim=-1, keepdim=True)
    frame_ids = batch[:, 8].long()
    tf_all = frame_transforms(params, c2w, cfg)
    tf = tf_all[frame_ids]                                            # [N,4,4]
    viewdirs_w = (tf[:, :3, :3] @ viewdirs[..., None])[..., 0]
    err = False
    if z_vals is None:
        with torch.no_grad():
            tf12 = tf_all[:, :3, :].reshape(-1, 12).detach().float().numpy()
            bnp = batch.detach().float().numpy()
            u, o, dw = rays_world_np(bnp, tf12)
            io = ray_trace_intervals(occ, o, dw)
            t_rand = None
            if t_rand_occ is not None:
                t_rand = np.concatenate([np.asarray(t_rand_occ, np.float32)] + ([np.asarray(t_rand_depth, np.float32)] if t_rand_depth is not None else []), -1)
            zv, err = sample_along_rays(io, u, bnp[:, 6], cfg, trunc, t_rand)
            z_vals = torch.from_numpy(zv)
    z_vals = z_vals.to(rays_d.dtype)
    S = z_vals.shape[1]
    pts = rays_d[:, None, :] * z_vals[:, :, None]                      # nerf_runner.py:1083
    x = (tf[:, None, :3, :3] @ pts[..., None])[..., 0] + tf[:, None, :3, 3]   # :1242-1243
    valid = (torch.abs(x) <= 1).all(dim=-1)                            # :1245
    xf = x.reshape(-1, 3)
    vf = valid.reshape(-1)
    E = (len(params['offsets']) - 1) * params['embeddings'].shape[1]
    emb = params['embeddings']
    if half:
        emb = emb.half().float()                                       # grid.py:50-51 (fp16 table under autocast)
    enc_valid = grid_encode((xf[vf] + 1) / 2, emb, params['offsets'], params['S'], params['H'],
                            exact_fma=False)
    if half:
        enc_valid = enc_valid.half().float()
    enc = torch.zeros(xf.shape[0], E, dtype=xf.dtype)
    enc = enc.index_put((vf.nonzero().reshape(-1),), enc_valid)
    views = sh_encode_deg3(viewdirs_w)                                 # :1282-1283
    if params.get('feature_data') is not None:                         # :1270-1278
        views = torch.cat([params['feature_data'][frame_ids], views], dim=-1)
    views_flat = views[:, None, :].expand(-1, S, -1).reshape(N * S, -1)
    raw = mlp_forward(params, enc, views_flat, half=half).reshape(N, S, 4)
    out = step_losses(raw, z_vals, valid, batch, trunc, cfg, params.get('pose_data'), params.get('feature_data'))
    if cfg.get('eikonal_weight', 0) > 0:                               # a15: intended maths of nerf_runner.py:734-738 (see eikonal_loss)
        out['eikonal_loss'] = eikonal_loss(params, xf, vf, cfg['eikonal_weight'], half=half)
        out['loss'] = out['loss'] + out['eikonal_loss']
    out.update(raw=raw, z_vals=z_vals, valid_samples=valid, x=x, tf=tf_all, sampling_error=err)
    return out


def adam_update(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-15):
    """torch.optim.Adam single-tensor math (nerf_runner.py:502). step is 1-based. In place."""
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


def lr_at(cfg, lr0, global_step):
    """schedule_lr, nerf_runner.py:579-583."""
    return lr0 * (cfg['decay_rate'] ** (float(global_step) / (cfg['n_step'] + 1)))


# ---------------------------------------------------------------------------------------------------------------------------
# Iso-surface extraction (checker for nof_marching_tets_*, include/nof.h; downstream of the path: extract_mesh,
# nerf_runner.py:1387-1404). Same algorithm in numpy: Kuhn split of each cell into 6 tetrahedra, one vertex per crossed grid
# edge interpolated from the lower-index end point (fp32, same operation order), triangles oriented towards increasing values.
MT_CORNER = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], np.int64)
MT_TET = np.array([[0, 5, 1, 6], [0, 1, 2, 6], [0, 2, 3, 6], [0, 3, 7, 6], [0, 7, 4, 6], [0, 4, 5, 6]], np.int64)


def marching_tets_np(field, iso=0.0):
    """field [nx,ny,nz] float32 -> (verts [T,3,3] float32 grid coordinates, keys [T,3] int64), cell-major / tet order."""
    f = np.asarray(field, np.float32)
    nx, ny, nz = f.shape
    iso = np.float32(iso)
    ci, cj, ck = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1), np.arange(nz - 1), indexing='ij')
    base = np.stack([ci.ravel(), cj.ravel(), ck.ravel()], -1)                       # [cells,3]
    verts_out, keys_out, order = [], [], []

    def edge_vertex(cells, qa, qb):
        pa = base[cells] + MT_CORNER[qa]
        pb = base[cells] + MT_CORNER[qb]
        ia = (pa[:, 0] * ny + pa[:, 1]) * nz + pa[:, 2]
        ib = (pb[:, 0] * ny + pb[:, 1]) * nz + pb[:, 2]
        a_lo = ia < ib
        lo = np.where(a_lo[:, None], pa, pb)
        hi = np.where(a_lo[:, None], pb, pa)
        vl = f[lo[:, 0], lo[:, 1], lo[:, 2]]
        vh = f[hi[:, 0], hi[:, 1], hi[:, 2]]
        t = ((iso - vl).astype(np.float32) / (vh - vl).astype(np.float32)).astype(np.float32)
        d = (hi - lo).astype(np.float32)
        # fmaf(t, d, lo) with d in {0,1}: t*d is exact, so one rounding like the fused form
        p = (t[:, None].astype(np.float64) * d.astype(np.float64) + lo.astype(np.float64)).astype(np.float32)
        key = np.where(a_lo, ia, ib) * 8 + ((hi - lo) @ np.array([4, 2, 1]))
        return p, key

    for t in range(6):
        q = MT_TET[t]
        pc = base[:, None, :] + MT_CORNER[q][None]                                  # [cells,4,3]
        v = f[pc[..., 0], pc[..., 1], pc[..., 2]]                                   # [cells,4]
        inside = v < iso
        n_in = inside.sum(1)
        for cells in [np.nonzero(n_in == 1)[0], np.nonzero(n_in == 3)[0], np.nonzero(n_in == 2)[0]]:
            if len(cells) == 0:
                continue
            ins = inside[cells]
            # positions (0..3) of inside / outside vertices, ascending like the kernel's loop
            idx = np.argsort(~ins, axis=1, kind='stable')                           # inside first, each group ascending
            ni = int(ins[0].sum())
            pin, pout = idx[:, :ni], idx[:, ni:]
            # winding from the tetrahedron's orientation (exact integer determinant), not from the triangle's geometry
            if ni == 1:
                L4 = np.concatenate([pin[:, :1], pout[:, :3]], 1)
            elif ni == 3:
                L4 = np.concatenate([pout[:, :1], pin[:, :3]], 1)
            else:
                L4 = np.concatenate([pin[:, :2], pout[:, :2]], 1)
            cpos = MT_CORNER[q]                                                     # [4,3] integer corner coordinates
            P = cpos[L4]                                                            # [cells,4,3]
            D = np.linalg.det((P[:, 1:] - P[:, :1]).astype(np.float64))
            flip = (D > 0.5) if ni == 3 else (D < -0.5)

            def ev(a, b):
                # a, b: per-cell positions inside the tet
                outp = np.zeros((len(cells), 3), np.float32)
                outk = np.zeros(len(cells), np.int64)
                for qa in range(4):
                    for qb in range(4):
                        m = (a == qa) & (b == qb)
                        if m.any():
                            p, k = edge_vertex(cells[m], q[qa], q[qb])
                            outp[m], outk[m] = p, k
                return outp, outk
            tris = []
            if ni == 1:
                tris.append((ev(pin[:, 0], pout[:, 0]), ev(pin[:, 0], pout[:, 1]), ev(pin[:, 0], pout[:, 2])))
            elif ni == 3:
                tris.append((ev(pout[:, 0], pin[:, 0]), ev(pout[:, 0], pin[:, 1]), ev(pout[:, 0], pin[:, 2])))
            else:
                p00, p01 = ev(pin[:, 0], pout[:, 0]), ev(pin[:, 0], pout[:, 1])
                p11, p10 = ev(pin[:, 1], pout[:, 1]), ev(pin[:, 1], pout[:, 0])
                tris.append((p00, p01, p11))
                tris.append((p00, p11, p10))
            for sub, (A, B, Cc) in enumerate(tris):
                Bp = np.where(flip[:, None], Cc[0], B[0]); Cp = np.where(flip[:, None], B[0], Cc[0])
                Bk = np.where(flip, Cc[1], B[1]); Ck = np.where(flip, B[1], Cc[1])
                verts_out.append(np.stack([A[0], Bp, Cp], 1))
                keys_out.append(np.stack([A[1], Bk, Ck], 1))
                order.append(np.stack([cells, np.full(len(cells), t), np.full(len(cells), sub)], 1))
    if not verts_out:
        return np.zeros((0, 3, 3), np.float32), np.zeros((0, 3), np.int64)
    V, K, O = np.concatenate(verts_out), np.concatenate(keys_out), np.concatenate(order)
    perm = np.lexsort((O[:, 2], O[:, 1], O[:, 0]))
    return V[perm], K[perm]


def weld_triangles(verts, keys):
    """(verts [T,3,3], keys [T,3]) -> (vertices [V,3], faces [F,3]) like bundlesdf_b200.ops.marching_tets."""
    uniq, first, inv = np.unique(keys.reshape(-1), return_index=True, return_inverse=True)
    vertices = verts.reshape(-1, 3)[first]
    faces = inv.reshape(-1, 3)
    keep = (faces[:, 0] != faces[:, 1]) & (faces[:, 1] != faces[:, 2]) & (faces[:, 0] != faces[:, 2])
    return vertices, faces[keep]


# ---------------------------------------------------------------------------------------------------------------------------
# a15 eikonal term — the INTENDED math of nerf_runner.py:734-738 + :1297-1302 (non-functional in the reference: train_loop renders
# with get_normals=False, and the normals branch uses grad_outputs=zeros / create_graph=False; the working variant is
# run_network_density :1342-1345 with grad_outputs=ones). This restatement EXTENDS the reference: normals are
# n = d sdf / d x with x the normalised point (what `inputs_flat` is there), the loss eikonal_weight * mean((|n|-1)^2) over the
# samples with sdf < 1, and its gradient w.r.t. the table and the SDF net comes from double backward (create_graph=True). Only
# valid samples have a network output (nerf_runner.py:1247), the others contribute sdf = 0 < 1 with n = 0, i.e. (0-1)^2 = 1,
# exactly like the reference's indexing `nerf_normals[sdf<1]` would.
def sdf_normals(params, x, valid, create_graph=False, half=False):
    """x [P,3] normalised points (requires no grad on entry), valid [P] bool. Returns (sdf [P], n [P,3]) with zeros at invalid
    samples; n is differentiable w.r.t. the parameters when create_graph=True. half: fp16 operand rounding of the AMP policy
    (table, encoding, weights, activations rounded to fp16, fp32 accumulation) — the casts are differentiable (identity)."""
    xr = x.detach().clone().requires_grad_(True)
    E = (len(params['offsets']) - 1) * params['embeddings'].shape[1]
    enc = torch.zeros(x.shape[0], E, dtype=x.dtype)
    idx = valid.nonzero().reshape(-1)
    emb = params['embeddings'].half().float() if half else params['embeddings']
    enc_valid = grid_encode((xr[idx] + 1) / 2, emb, params['offsets'], params['S'], params['H'], exact_fma=False)
    if half:
        enc_valid = enc_valid.half().float()
    enc = enc.index_put((idx,), enc_valid)
    sdf_v = mlp_forward_sdf(params, enc[idx], half=half)
    sdf = torch.zeros(x.shape[0], dtype=x.dtype).index_put((idx,), sdf_v)
    (n,) = torch.autograd.grad(sdf_v.sum(), xr, create_graph=create_graph, allow_unused=True)
    if n is None:
        n = torch.zeros_like(xr)
    return sdf, n


def eikonal_loss(params, x, valid, eikonal_weight, half=False):
    """eikonal_weight * mean over {sdf < 1} of (|n| - 1)^2, differentiable w.r.t. params['embeddings'] and the sigma_net weights."""
    sdf, n = sdf_normals(params, x, valid, create_graph=True, half=half)
    sel = sdf.detach() < 1
    if not bool(sel.any()):
        return torch.zeros((), dtype=x.dtype)
    return ((torch.linalg.norm(n[sel], dim=-1) - 1) ** 2).mean() * eikonal_weight
