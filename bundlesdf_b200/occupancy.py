"""Dense occupancy replacement of the reference's kaolin-backed OctreeManager (Utils.py:359-475).

The reference ray-traces a sparse octree at level floor(log2(2/(0.02*sc))) = 3..4 for hand-held objects, i.e. an 8^3..16^3
cell grid. At that size a dense bitmask (512 B for 16^3) staged in shared memory beats any tree: the fused sampler
(nof_ray_march) walks it with a voxel DDA, one warp per ray, with no host synchronisation (the reference's ray_trace
does a .item() per call, Utils.py:467). The build rule is the reference's (nerf_runner.py:436-476): quantise the cloud
at max_level, dilate by the 27-neighbourhood, clip the centres, quantise again (kaolin quantize_points) and mark a
ray-tracing-level cell occupied iff any descendant is.
"""
import numpy as np
import torch

from . import ops


def octree_levels(cfg):
    sc = cfg['sc_factor']
    max_level = int(np.ceil(np.log2(2.0 / (cfg['octree_smallest_voxel_size'] * sc))))      # nerf_runner.py:444-447
    level = int(np.floor(np.log2(2.0 / (cfg['octree_raytracing_voxel_size'] * sc))))       # nerf_runner.py:1058-1059
    return max_level, level


class OctreeManager:
    def __init__(self, pts=None, max_level=None, octree=None, level=None, device=None):
        """pts: [M,3] tensor of (dilated) voxel centres in [-1,1] at max_level, like the reference passes (nerf_runner.py:476);
        octree: a dict previously returned by `.octree` (checkpoint round trip, nerf_runner.py:542-543)."""
        if octree is not None:
            self.max_level = int(octree['max_level'])
            self.level = int(octree['level'])
            occ = torch.as_tensor(octree['occ']).bool()
        else:
            assert level is not None and level <= max_level
            self.max_level, self.level = int(max_level), int(level)
            n_max = 2 ** self.max_level
            q = torch.clamp(torch.floor((pts.double() + 1) / 2 * n_max), 0, n_max - 1).long()     # kaolin quantize_points
            q = q >> (self.max_level - self.level)
            n = 2 ** self.level
            occ = torch.zeros(n, n, n, dtype=torch.bool, device=pts.device)
            occ[q[:, 0], q[:, 1], q[:, 2]] = True
        self.device = device or (pts.device if pts is not None else torch.device('cpu'))
        self.occ = occ.to(self.device)
        self.n = 2 ** self.level
        self.occ_bits = ops.pack_occupancy(self.occ.cpu().numpy()).to(self.device)
        self.n_vox = int(self.occ.sum().item())

    @property
    def octree(self):
        """Serializable state (stands in for the kaolin octree byte tensor stored by save_weights, nerf_runner.py:565-566)."""
        return {'occ': self.occ.cpu(), 'level': self.level, 'max_level': self.max_level}

    def get_center_ids(self, x, level=None):
        """Utils.py:392-394 stand-in: index (>=0) of the ray-tracing-level cell containing x if it is occupied, else -1."""
        n = self.n
        q = torch.clamp(torch.floor((x + 1) / 2 * n), 0, n - 1).long()
        lin = (q[:, 0] * n + q[:, 1]) * n + q[:, 2]
        inside = ((x >= -1) & (x <= 1)).all(-1)
        ok = self.occ.reshape(-1)[lin] & inside
        return torch.where(ok, lin, torch.full_like(lin, -1))

    def ray_trace(self, rays_o, rays_d, level=None, debug=False):
        """Same contract as Utils.py:443-475: (rays_near [N,1], rays_far [N,1], rays_pid [N,1] (-1: unused), intervals [N,I,2])
        with I = 3*2^level fixed (no host sync to find the batch maximum). rays_d must be unit length."""
        assert level is None or level == self.level, 'the occupancy is stored at the ray-tracing level only'
        N = rays_o.shape[0]
        dev = rays_o.device
        rows = torch.zeros(N, 12, device=dev)
        rows[:, 0:3] = rays_d
        rows[:, 8] = torch.arange(N, device=dev, dtype=torch.float32)
        assert N < (1 << 24), 'per-ray frames are indexed through a float32 column'
        tf = torch.zeros(N, 12, device=dev)
        tf[:, 0] = 1; tf[:, 5] = 1; tf[:, 10] = 1
        tf[:, 3] = rays_o[:, 0]; tf[:, 7] = rays_o[:, 1]; tf[:, 11] = rays_o[:, 2]
        _, inter = ops.ray_march(rows.contiguous(), tf.contiguous(), self.occ_bits, self.level, 1, 0, 0.0, 1.0, 0.0, 1.0, t_rand=None,
                                 perturb=False, want_intervals=True)
        near = inter[:, 0, 0].reshape(-1, 1)
        far = inter[:, :, 1].max(dim=-1)[0].reshape(-1, 1)
        pid = -torch.ones_like(near)
        return near, far, pid, inter


def build_occupancy_points(pts, cfg):
    """nerf_runner.py:443-465: dilated voxel centres at max_level from the raw cloud (tensor [M,3] in [-1,1])."""
    max_level, level = octree_levels(cfg)
    vox = 2.0 / (2 ** max_level)
    dilate_radius = max(1, int(np.ceil(cfg['octree_dilate_size'] / cfg['octree_smallest_voxel_size'])))
    coords = torch.floor((pts.float() + 1) / vox).long()
    coords = torch.unique(coords, dim=0)
    r = torch.tensor([-1, 0, 1], device=pts.device)
    shifts = torch.stack(torch.meshgrid(r, r, r, indexing='ij'), -1).reshape(-1, 3)
    for _ in range(dilate_radius):
        coords = torch.unique((coords[None] + shifts[:, None]).reshape(-1, 3), dim=0)
    centers = torch.clip((coords.double() + 0.5) * vox - 1, -1, 1)
    return centers, max_level, level
