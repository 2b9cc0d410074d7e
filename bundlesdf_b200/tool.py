"""Scene normalisation and the on-disk formats either side of the Neural-Object-Field trainer.

What `bundlesdf.py` does around `NerfRunner` and what `tool.py` of the reference computes, with the same function names and results, on
numpy / scipy / scikit-learn only (the reference goes through Open3D, which is third-party and absent here):

  compute_translation_scales   tool.py:28-39   translation = -bbox centre, sc_factor = 0.9 * max_dim / largest extent (biggest DBSCAN cluster)
  compute_scene_bounds         tool.py:67-132  fuse the masked depth maps into one cloud in the object frame, voxel down-sample, normalise
  the files of a BundleSDF run (bundlesdf.py:640-735): cam_K.txt, ob_in_cam/<id>.txt, <stamp>/keyframes.yml, nerf/config.yml,
  normalization.yml, trainval_poses.txt — readers and writers that round-trip with the reference's own yaml / numpy calls.

The Open3D pieces are restated, not linked: `voxel_down_sample` = mean of the points of each occupied voxel (voxel index = floor((p - min) /
size), Open3D's rule), `remove_statistical_outlier(nb_neighbors, std_ratio)` = drop points whose mean distance to their nb_neighbors nearest
neighbours exceeds mean + std_ratio * std of that statistic. PARITY UNPINNED against Open3D itself (not installable offline); pinned are the
closed-form parts (tests/test_tool_formats.py against the reference's own tool.compute_translation_scales)."""
import glob
import os

import numpy as np
import yaml

glcam_in_cvcam = np.array([[1, 0, 0, 0], [0, -1, 0, 0], [0, 0, -1, 0], [0, 0, 0, 1]], dtype=np.float64)     # Utils.py:37-40


# ------------------------------------------------------------------------------------------------ point-cloud helpers
def depth2xyzmap(depth, K):
    """Utils.py:219-231: back-project a depth image (metres) to camera-frame points [H,W,3] float32; depth < 0.1 -> 0."""
    H, W = depth.shape[:2]
    vs, us = np.meshgrid(np.arange(H), np.arange(W), sparse=False, indexing='ij')
    zs = depth
    xs = (us - K[0, 2]) * zs / K[0, 0]
    ys = (vs - K[1, 2]) * zs / K[1, 1]
    pts = np.stack((xs, ys, zs), axis=-1).astype(np.float32)
    pts[depth < 0.1] = 0
    return pts


def voxel_down_sample(pts, voxel_size, colors=None):
    """Open3D PointCloud.voxel_down_sample: one point per occupied voxel = the mean of the points (and colours) that fall into it."""
    pts = np.asarray(pts, dtype=np.float64)
    if len(pts) == 0:
        return (pts, colors) if colors is not None else pts
    lo = pts.min(axis=0) - voxel_size * 0.5
    key = np.floor((pts - lo) / voxel_size).astype(np.int64)
    _, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    inv = inv.reshape(-1)
    out = np.zeros((len(cnt), 3))
    np.add.at(out, inv, pts)
    out /= cnt[:, None]
    if colors is None:
        return out
    col = np.zeros((len(cnt), colors.shape[1]))
    np.add.at(col, inv, np.asarray(colors, dtype=np.float64))
    return out, col / cnt[:, None]


def remove_statistical_outlier(pts, nb_neighbors=30, std_ratio=2.0):
    """Open3D remove_statistical_outlier: keep mask of the points whose mean k-NN distance is <= mean + std_ratio * std."""
    from scipy.spatial import cKDTree
    pts = np.asarray(pts, dtype=np.float64)
    if len(pts) <= 1:
        return np.ones(len(pts), dtype=bool)
    k = min(nb_neighbors, len(pts))
    d, _ = cKDTree(pts).query(pts, k=k)
    mean_d = d.reshape(len(pts), -1).mean(axis=1)           # Open3D averages over the k neighbours INCLUDING the point itself (distance 0)
    return mean_d <= mean_d.mean() + std_ratio * mean_d.std()


def find_biggest_cluster(pts, eps=0.06, min_samples=1):
    """tool.py:18-25."""
    from sklearn.cluster import DBSCAN
    db = DBSCAN(eps=eps, min_samples=min_samples, n_jobs=-1)
    db.fit(pts)
    ids, cnts = np.unique(db.labels_, return_counts=True)
    best = ids[cnts.argsort()[-1]]
    keep = db.labels_ == best
    return pts[keep], keep


def compute_translation_scales(pts, max_dim=2, cluster=True, eps=0.06, min_samples=1):
    """tool.py:28-39: (translation_cvcam, sc_factor, keep_mask)."""
    if cluster:
        pts, keep = find_biggest_cluster(pts, eps, min_samples)
    else:
        keep = np.ones(len(pts), dtype=bool)
    hi, lo = pts.max(axis=0), pts.min(axis=0)
    center = (hi + lo) / 2
    sc_factor = max_dim / (hi - lo).max()                   # normalise to [-1, 1]
    sc_factor *= 0.9                                        # reserve some space
    return -center, sc_factor, keep


def compute_scene_bounds_worker(K, glcam_in_world, use_mask, rgb, depth, mask):
    """tool.py:42-64 for in-memory frames: masked back-projection, 1 cm voxel grid, statistical outlier removal, to the world frame."""
    xyz = depth2xyzmap(depth, K)
    valid = depth >= 0.1
    if use_mask:
        valid = valid & (np.asarray(mask).reshape(depth.shape) > 0)
    pts = xyz[valid].reshape(-1, 3)
    if len(pts) == 0:
        return None
    colors = np.asarray(rgb)[valid].reshape(-1, 3) / 255.0
    pts, colors = voxel_down_sample(pts, 0.01, colors)
    keep = remove_statistical_outlier(pts, nb_neighbors=30, std_ratio=2.0)
    pts, colors = pts[keep], colors[keep]
    cam_in_world = glcam_in_world @ glcam_in_cvcam
    return pts @ cam_in_world[:3, :3].T + cam_in_world[:3, 3], colors


def compute_scene_bounds(glcam_in_worlds, K, rgbs, depths, masks, use_mask=True, base_dir=None, cluster=True, translation_cvcam=None, sc_factor=None,
                         eps=0.06, min_samples=1):
    """tool.py:67-132 (in-memory branch used by bundlesdf.py:186,699). Returns (sc_factor, translation_cvcam, pts_real_scale, pts_normalized)
    — the last two are [M,3] arrays where the reference returns Open3D clouds; wrap them in `synthetic.PointCloud` for NerfRunner."""
    clouds = []
    for i in range(len(rgbs)):
        r = compute_scene_bounds_worker(K, glcam_in_worlds[i], use_mask, rgbs[i], depths[i], masks[i])
        if r is not None:
            clouds.append(r[0])
    pts = voxel_down_sample(np.concatenate(clouds, 0), eps / 5)
    if translation_cvcam is None:
        translation_cvcam, sc_factor, keep = compute_translation_scales(pts, cluster=cluster, eps=eps, min_samples=min_samples)
    else:
        translation_cvcam = np.asarray(translation_cvcam, dtype=np.float64)
        keep = (np.abs((pts + translation_cvcam) * sc_factor) < 1).all(axis=-1)
    pts_real = pts[keep]
    if base_dir is not None:
        write_normalization(os.path.join(base_dir, 'normalization.yml'), translation_cvcam, sc_factor)
    return float(sc_factor), translation_cvcam, pts_real, (pts_real + translation_cvcam) * sc_factor


# ------------------------------------------------------------------------------------------------ files of a BundleSDF run
def write_normalization(path, translation_cvcam, sc_factor):
    """tool.py:123-128."""
    os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
    with open(path, 'w') as f:
        yaml.dump({'translation_cvcam': np.asarray(translation_cvcam).tolist(), 'sc_factor': float(sc_factor)}, f)


def write_nerf_config(path, cfg):
    """bundlesdf.py:211-214 / 714-717: the NeRF cfg with numpy values converted, as `nerf/config.yml` (run_global_nerf re-reads
    sc_factor and translation from the newest one, bundlesdf.py:692-697)."""
    os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
    out = {}
    for k, v in cfg.items():
        out[k] = v.tolist() if isinstance(v, np.ndarray) else (float(v) if isinstance(v, np.floating) else (int(v) if isinstance(v, np.integer) else v))
    with open(path, 'w') as f:
        yaml.dump(out, f)


def read_normalization_from_run(debug_dir):
    """bundlesdf.py:692-697: (sc_factor, translation) of the newest **/nerf/config.yml under a run directory, or (None, None)."""
    files = sorted(glob.glob(f'{debug_dir}/**/nerf/config.yml', recursive=True))
    if not files:
        return None, None
    tmp = yaml.safe_load(open(files[-1], 'r'))
    return float(tmp['sc_factor']), np.array(tmp['translation'], dtype=np.float64)


def write_keyframes(path, frame_ids, cam_in_obs):
    """<stamp>/keyframes.yml as the tracker writes it (read at bundlesdf.py:645-660): {keyframe_<id>: {cam_in_ob: [16 floats]}}."""
    os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
    data = {f'keyframe_{fid}': {'cam_in_ob': np.asarray(T, dtype=np.float64).reshape(-1).tolist()} for fid, T in zip(frame_ids, cam_in_obs)}
    with open(path, 'w') as f:
        yaml.dump(data, f, sort_keys=False)


def read_keyframes(path):
    """bundlesdf.py:645-660: (frame_ids [str], cam_in_obs [K,4,4]) in file order."""
    kf = yaml.safe_load(open(path, 'r'))
    keys = list(kf.keys())
    return [k.replace('keyframe_', '') for k in keys], np.array([np.array(kf[k]['cam_in_ob']).reshape(4, 4) for k in keys])


def write_pose(path, T):
    """ob_in_cam/<id>.txt, poses_after_nerf.txt, trainval_poses.txt: np.savetxt of the 4x4 rows (bundlesdf.py:555,711,742)."""
    os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
    np.savetxt(path, np.asarray(T, dtype=np.float64).reshape(-1, 4))


def read_pose(path):
    return np.loadtxt(path).reshape(-1, 4, 4) if np.loadtxt(path).size > 16 else np.loadtxt(path).reshape(4, 4)


def last_stamp(debug_dir):
    """bundlesdf.py:642-643."""
    files = sorted(glob.glob(f'{debug_dir}/ob_in_cam/*'))
    return os.path.basename(files[-1]).replace('.txt', '') if files else None


def load_global_refine_inputs(debug_dir):
    """The part of run_global_nerf (bundlesdf.py:640-660) that only reads files: K, keyframe ids and their GL camera poses in the object frame."""
    K = np.loadtxt(f'{debug_dir}/cam_K.txt').reshape(3, 3)
    stamp = last_stamp(debug_dir)
    ids, cam_in_obs = read_keyframes(f'{debug_dir}/{stamp}/keyframes.yml')
    return K, ids, cam_in_obs @ glcam_in_cvcam
