"""Deterministic synthetic RGBD sequences shaped like BundleSDF's inputs: an analytic-SDF "milk jug"
(rounded box + cylinder neck + torus handle) sphere-traced from an orbiting 640x480 pinhole camera. Produces exactly
the constructor inputs of NerfRunner (nerf_runner.py:112: images, depths, masks, poses, K, build_octree_pcd) after the
reference's preprocess_data (nerf_helpers.py:218-240) and scene normalisation (tool.py:28-39, bundlesdf.py:151).
There is no network in the build/bench environment, so this replaces the datasets (YCBInEOAT / HO3D)."""
import math

import numpy as np
import torch

BAD_DEPTH = 99.0          # Utils.py:34
GLCAM_IN_CVCAM = np.array([[1, 0, 0, 0], [0, -1, 0, 0], [0, 0, -1, 0], [0, 0, 0, 1]], dtype=np.float64)   # Utils.py:37-40


def jug_sdf(p):
    """Signed distance (metres) of the object, p [...,3] torch."""
    x, y, z = p[..., 0], p[..., 1], p[..., 2]
    # rounded box body 0.10 x 0.10 x 0.18, corner radius 0.015
    r = 0.015
    q = torch.stack([x.abs() - (0.05 - r), y.abs() - (0.05 - r), z.abs() - (0.09 - r)], -1)
    body = torch.linalg.norm(q.clamp(min=0), dim=-1) + q.max(dim=-1).values.clamp(max=0) - r
    # neck: cylinder r=0.025, z in [0.09, 0.14]
    dxy = torch.sqrt(x * x + y * y) - 0.025
    dz = (z - 0.115).abs() - 0.025
    neck = torch.sqrt(dxy.clamp(min=0) ** 2 + dz.clamp(min=0) ** 2) + torch.maximum(dxy, dz).clamp(max=0)
    # handle: torus in the x-z plane centred at (0.05, 0, 0.02), R=0.035, r=0.009
    qx = torch.sqrt((x - 0.05) ** 2 + (z - 0.02) ** 2) - 0.035
    handle = torch.sqrt(qx * qx + y * y) - 0.009
    return torch.minimum(torch.minimum(body, neck), handle)


def jug_albedo(p):
    """Smooth procedural colour of a surface point (learnable by a small MLP)."""
    x, y, z = p[..., 0], p[..., 1], p[..., 2]
    r = 0.55 + 0.40 * torch.sin(25.0 * x + 1.0) * torch.cos(18.0 * z)
    g = 0.50 + 0.35 * torch.sin(22.0 * y + 14.0 * z)
    b = 0.45 + 0.40 * torch.cos(20.0 * x - 16.0 * y + 0.5)
    return torch.stack([r, g, b], -1).clamp(0.02, 0.98)


def orbit_pose_cv(i, n, radius=0.6):
    """cam-in-object (OpenCV convention: x right, y down, z forward), looking at the origin."""
    az = 2.0 * math.pi * i / max(n, 1)
    el = math.radians(20.0 + 10.0 * math.sin(2.0 * az))
    c = np.array([radius * math.cos(el) * math.cos(az), radius * math.cos(el) * math.sin(az), radius * math.sin(el)])
    fwd = -c / np.linalg.norm(c)
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up); right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    T = np.eye(4)
    T[:3, 0], T[:3, 1], T[:3, 2], T[:3, 3] = right, down, fwd, c
    return T


@torch.no_grad()
def render_frame(cam_in_ob_cv, K, H, W, device, n_steps=48):
    """Sphere-trace the SDF. Returns rgb [H,W,3] float 0..1, depth [H,W] metres (0 = miss), mask [H,W] bool."""
    T = torch.as_tensor(cam_in_ob_cv, dtype=torch.float32, device=device)
    v, u = torch.meshgrid(torch.arange(H, device=device, dtype=torch.float32), torch.arange(W, device=device, dtype=torch.float32), indexing='ij')
    d_cam = torch.stack([(u - K[0, 2]) / K[0, 0], (v - K[1, 2]) / K[1, 1], torch.ones_like(u)], -1)     # z-depth parametrisation
    nrm = d_cam.norm(dim=-1, keepdim=True)
    d_w = (d_cam / nrm) @ T[:3, :3].T
    o = T[:3, 3]
    t = torch.full((H, W), 0.3, device=device)
    for _ in range(n_steps):
        t = t + jug_sdf(o + d_w * t[..., None])
        t = t.clamp(max=1.5)
    p = o + d_w * t[..., None]
    hit = (jug_sdf(p).abs() < 2e-4) & (t < 1.4)
    depth = torch.where(hit, t / nrm[..., 0], torch.zeros_like(t))                                      # z-depth
    rgb = torch.where(hit[..., None], jug_albedo(p), torch.zeros_like(p))
    return rgb, depth, hit


def make_sequence(n_frames, H=480, W=640, device='cpu', seed=0, pose_noise=False, depth_noise_m=0.001, frame_stride=1,
                  total_frames=None, sc_factor=None):
    """Returns dict with the reference NerfRunner's constructor inputs (already preprocessed / normalised):
      images [F,H,W,3] float32 0..1, depths [F,H,W,1] float32 (metres*sc_factor, BAD_DEPTH*sc where invalid),
      masks [F,H,W,1] uint8, poses [F,4,4] float64 (GL cam-in-object, normalised), K [3,3], pcd_normalized [M,3],
      sc_factor, translation, poses_gt (without the injected noise).
    pose_noise: translation N(0,5mm), rotation N(0,2deg) on every frame but 0 (pose-refinement configs)."""
    rng = np.random.default_rng(seed)
    K = np.array([[600.0 * W / 640, 0, W / 2.0], [0, 600.0 * H / 480, H / 2.0], [0, 0, 1]], dtype=np.float64)
    Kt = torch.as_tensor(K, dtype=torch.float32, device=device)
    total = total_frames or n_frames * frame_stride
    images, depths, masks, poses_cv = [], [], [], []
    clouds = []
    for k in range(n_frames):
        T = orbit_pose_cv(k * frame_stride, total)
        rgb, depth, hit = render_frame(T, Kt, H, W, device)
        if depth_noise_m > 0:
            noise = torch.from_numpy(rng.normal(0, depth_noise_m, size=(H, W)).astype(np.float32)).to(device)
            depth = torch.where(hit, depth + noise, depth)
        images.append(rgb.cpu().numpy()); depths.append(depth.cpu().numpy()); masks.append(hit.cpu().numpy()); poses_cv.append(T)
        # back-projected cloud (object frame) for scene bounds / octree, subsampled
        vs, us = np.nonzero(masks[-1][::4, ::4])
        z = depths[-1][::4, ::4][vs, us]
        pc = np.stack([(us * 4 - K[0, 2]) / K[0, 0] * z, (vs * 4 - K[1, 2]) / K[1, 1] * z, z], -1)
        clouds.append((T[:3, :3] @ pc.T).T + T[:3, 3])
    cloud = np.concatenate(clouds, 0)
    # tool.py:28-39 + bundlesdf.py:151
    lo, hi = cloud.min(0), cloud.max(0)
    translation = -(lo + hi) / 2.0
    if sc_factor is None:
        sc_factor = 0.9 * 2.0 / float((hi - lo).max()) * 0.7
    poses_cv = np.stack(poses_cv)
    poses_gl_gt = poses_cv @ GLCAM_IN_CVCAM                      # bundlesdf.py:145
    poses_gl = poses_gl_gt.copy()
    if pose_noise:
        for i in range(1, n_frames):
            w = rng.normal(0, math.radians(2.0), 3)
            th = np.linalg.norm(w)
            Kx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
            R = np.eye(3) + math.sin(th) / th * Kx + (1 - math.cos(th)) / th ** 2 * Kx @ Kx
            D = np.eye(4); D[:3, :3] = R; D[:3, 3] = rng.normal(0, 0.005, 3)
            poses_gl[i] = D @ poses_gl[i]
    images = np.stack(images).astype(np.float32)
    depths = np.stack(depths).astype(np.float32)
    masks = np.stack(masks)
    # preprocess_data (nerf_helpers.py:218-240)
    depths[depths < 0.1] = BAD_DEPTH
    depths[~masks] = BAD_DEPTH
    images[~masks] = 128.0 / 255.0                               # BAD_COLOR
    depths = (depths * sc_factor)[..., None]
    def norm(P):
        P = P.copy(); P[:, :3, 3] += translation; P[:, :3, 3] *= sc_factor; return P
    pcd = (cloud + translation) * sc_factor
    # voxel downsample 1 cm (in normalised units)
    vox = 0.01 * sc_factor
    key = np.floor(pcd / vox).astype(np.int64)
    _, first = np.unique(key, axis=0, return_index=True)
    pcd = pcd[np.sort(first)]
    return dict(images=images, depths=depths, masks=masks[..., None].astype(np.uint8), poses=norm(poses_gl), poses_gt=norm(poses_gl_gt),
                K=K, pcd_normalized=pcd.astype(np.float64), sc_factor=float(sc_factor), translation=translation, H=H, W=W)


class PointCloud:
    """Minimal stand-in for the Open3D cloud the reference passes as build_octree_pcd (only `.points` and
    `.voxel_down_sample` are used: nerf_runner.py:127,376)."""

    def __init__(self, points):
        self.points = np.asarray(points, dtype=np.float64)

    def voxel_down_sample(self, voxel_size):
        key = np.floor(self.points / voxel_size).astype(np.int64)
        _, first = np.unique(key, axis=0, return_index=True)
        return PointCloud(self.points[np.sort(first)])


def default_cfg(**over):
    """The hot-path keys of the reference's config.yml (values copied as documented in SURVEY.md §5; config.yml:2-93)."""
    cfg = dict(n_step=500, N_rand=2048, lrate=0.01, lrate_pose=0.01, decay_rate=0.1, chunk=99999999999, netchunk=6553600, amp=True,
               N_samples=128, N_samples_around_depth=64, N_importance=0, perturb=1, use_viewdirs=1, i_embed=1, i_embed_views=2,
               multires=8, multires_views=3, feature_grid_dim=2, raw_noise_std=0, i_print=999999, i_img=999999, i_weights=999999,
               i_mesh=999999, i_pose=999999, save_octree_clouds=False, finest_res=128, base_res=16, num_levels=4, log2_hashmap_size=22,
               n_train_image=300, use_octree=1, first_frame_weight=10, denoise_depth_use_octree_cloud=True,
               octree_smallest_voxel_size=0.02, octree_raytracing_voxel_size=0.02, octree_dilate_size=0.02, down_scale_ratio=1,
               bounding_box=[[-1, -1, -1], [1, 1, 1]], use_mask=1, dilate_mask_size=0, rays_valid_depth_only=True, near=0.1, far=2,
               rgb_weight=10, depth_weight=0, trunc=0.01, trunc_start=0.01, sdf_lambda=5, neg_trunc_ratio=1, trunc_decay_type='',
               fs_weight=100, empty_weight=0.01, fs_rgb_weight=0, trunc_weight=6000, tv_loss_weight=0, frame_features=0,
               optimize_poses=1, pose_reg_weight=0, eikonal_weight=0, feature_reg_weight=0.1, share_coarse_fine=1, mode='sdf',
               fs_sdf=0.001, mesh_resolution=0.005, max_trans=0.02, max_rot=20, no_batching=0, save_dir='/tmp/nof_out',
               sc_factor=1.0, translation=[0, 0, 0])
    cfg.update(over)
    return cfg
