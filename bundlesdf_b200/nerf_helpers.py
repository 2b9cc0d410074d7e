"""Model pieces of the Neural Object Field with the reference's class names, constructor arguments and state_dict
keys (nerf_helpers.py of the reference), so checkpoints and callers are interchangeable. The modules are *containers*:
during training their parameters alias flat device buffers that libnof_sm100's fused kernels read and update in
place; `forward` methods exist for API parity (evaluation, mesh extraction) and run native kernels where one exists.
"""
import numpy as np
import torch
import torch.nn as nn

from . import ops

BAD_DEPTH = 99        # Utils.py:34
BAD_COLOR = 128       # Utils.py:35


class SHEncoder(nn.Module):
    """Degree<=3 real spherical harmonics of a unit direction (reference nerf_helpers.py:22-105). The fused step kernel
    evaluates the same 9 polynomials per ray on device; this module is the eval-time mirror (parameter free)."""

    def __init__(self, input_dim=3, degree=3):
        super().__init__()
        assert input_dim == 3 and 1 <= degree <= 3, 'only degree<=3 (multires_views: 3) is built'
        self.input_dim, self.degree, self.out_dim = input_dim, degree, degree ** 2

    def forward(self, d, **kwargs):
        x, y, z = d.unbind(-1)
        c = [torch.full_like(x, 0.28209479177387814)]
        if self.degree > 1:
            c += [-0.4886025119029199 * y, 0.4886025119029199 * z, -0.4886025119029199 * x]
        if self.degree > 2:
            c += [1.0925484305920792 * x * y, -1.0925484305920792 * y * z, 0.31539156525252005 * (2.0 * z * z - x * x - y * y),
                  -1.0925484305920792 * x * z, 0.5462742152960396 * (x * x - y * y)]
        return torch.stack(c, -1)


class FeatureArray(nn.Module):
    """Per-frame latent code, N(0,1) init (reference nerf_helpers.py:108-124)."""

    def __init__(self, num_frames, num_channels):
        super().__init__()
        self.num_frames, self.num_channels = num_frames, num_channels
        self.data = nn.parameter.Parameter(torch.normal(0, 1, size=[num_frames, num_channels]).float(), requires_grad=True)

    def __call__(self, ids):
        return self.data[ids]


class PoseArray(nn.Module):
    """Per-frame se(3) correction, zero init, frame 0 pinned to identity (reference nerf_helpers.py:127-154)."""

    def __init__(self, num_frames, max_trans, max_rot):
        super().__init__()
        self.num_frames, self.max_trans, self.max_rot = num_frames, max_trans, max_rot
        self.data = nn.parameter.Parameter(torch.zeros([num_frames, 6]).float(), requires_grad=True)

    def get_matrices(self, ids):
        """[len(ids),4,4] correction matrices, evaluated by the native pose kernel (identity c2w)."""
        if not torch.is_tensor(ids):
            ids = torch.as_tensor(np.asarray(ids)).long()
        dev = self.data.device
        eye = torch.eye(4, device=dev).repeat(self.num_frames, 1, 1).contiguous()
        tf = ops.pose_forward(self.data.detach().contiguous(), eye, self.max_trans, self.max_rot)
        Ts = eye.clone()
        Ts[:, :3, :] = tf.view(-1, 3, 4)
        return Ts[ids.to(dev).long().reshape(-1)]


class NeRFSmall(nn.Module):
    """SDF net E->64->16 (+0.1 bias on the last layer) and colour net (views+15)->64->64->3, same Sequential layout and
    therefore the same state_dict keys as the reference (nerf_helpers.py:243-321)."""

    def __init__(self, num_layers=3, hidden_dim=64, geo_feat_dim=15, num_layers_color=4, hidden_dim_color=64, input_ch=3, input_ch_views=3):
        super().__init__()
        self.input_ch, self.input_ch_views = input_ch, input_ch_views
        self.num_layers, self.hidden_dim, self.geo_feat_dim = num_layers, hidden_dim, geo_feat_dim
        self.num_layers_color, self.hidden_dim_color = num_layers_color, hidden_dim_color
        sigma = []
        for l in range(num_layers):
            i = input_ch if l == 0 else hidden_dim
            o = 1 + geo_feat_dim if l == num_layers - 1 else hidden_dim
            sigma.append(nn.Linear(i, o, bias=True))
            if l != num_layers - 1:
                sigma.append(nn.ReLU(inplace=True))
        self.sigma_net = nn.Sequential(*sigma)
        torch.nn.init.constant_(self.sigma_net[-1].bias, 0.1)
        color = []
        for l in range(num_layers_color):
            i = input_ch_views + geo_feat_dim if l == 0 else hidden_dim
            o = 3 if l == num_layers_color - 1 else hidden_dim
            color.append(nn.Linear(i, o, bias=True))
            if l != num_layers_color - 1:
                color.append(nn.ReLU(inplace=True))
        self.color_net = nn.Sequential(*color)

    def is_fused_layout(self):
        return (self.num_layers == 2 and self.hidden_dim == 64 and self.geo_feat_dim == 15 and self.num_layers_color == 3
                and self.hidden_dim_color == 64)

    def forward_sdf(self, x):
        return self.sigma_net(x)[..., 0]

    def forward(self, x):
        x = x.float()
        pts, views = torch.split(x, [self.input_ch, self.input_ch_views], dim=-1)
        h = self.sigma_net(pts)
        sigma, geo = h[..., 0], h[..., 1:]
        color = self.color_net(torch.cat([views, geo], dim=-1))
        return torch.cat([color, sigma.unsqueeze(-1)], -1)


def get_embedder(multires, cfg, i=0, octree_m=None):
    """Reference nerf_helpers.py:191-215; only the shipped encoders are built (i_embed: 1 hash grid, i_embed_views: 2 SH)."""
    if i == -1:
        return nn.Identity(), 3
    if i == 1:
        from .mycuda.torch_ngp_grid_encoder.grid import GridEncoder
        embed = GridEncoder(input_dim=3, n_levels=cfg['num_levels'], log2_hashmap_size=cfg['log2_hashmap_size'],
                            desired_resolution=cfg['finest_res'], base_resolution=cfg['base_res'], level_dim=cfg['feature_grid_dim'])
        return embed, embed.out_dim
    if i == 2:
        embed = SHEncoder(degree=cfg['multires_views'])
        return embed, embed.out_dim
    raise NotImplementedError(f'embedder i={i}: the reference config ships i_embed=1 / i_embed_views=2; the frequency embedder '
                              f'(i=0) and the octree grid (i=3) are dead code under config.yml and are not built')


def get_camera_rays_np(H, W, K):
    """Reference nerf_helpers.py:358-363 (OpenGL camera: x right, y up, z backward)."""
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing='xy')
    return np.stack([(i - K[0, 2]) / K[0, 0], -(j - K[1, 2]) / K[1, 1], -np.ones_like(i)], axis=-1)


def preprocess_data(rgbs, depths, masks, normal_maps, poses, sc_factor, translation):
    """Reference nerf_helpers.py:218-240 (in place on the inputs, like the reference)."""
    depths[depths < 0.1] = BAD_DEPTH
    if masks is not None:
        rgbs[masks == 0] = BAD_COLOR
        depths[masks == 0] = BAD_DEPTH
        if normal_maps is not None:
            normal_maps[..., [1, 2]] *= -1
            normal_maps[masks == 0] = 0
        masks = masks[..., None]
    rgbs = (rgbs / 255.0).astype(np.float32)
    depths *= sc_factor
    depths = depths[..., None]
    poses[:, :3, 3] += translation
    poses[:, :3, 3] *= sc_factor
    return rgbs, depths, masks, normal_maps, poses
