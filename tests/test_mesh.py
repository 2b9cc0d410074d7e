"""Iso-surface extraction: the numpy restatement's properties on CPU, the CUDA kernels against it on the GPU."""
import numpy as np
import pytest
import torch

from bundlesdf_b200.mesh import TriMesh
from oracle import nof_oracle as O


def _grid(n):
    ax = np.linspace(-1, 1, n).astype(np.float32)
    return np.meshgrid(ax, ax, ax, indexing='ij')


def _sphere(n, r=0.6, c=(0.07, -0.03, 0.11)):
    X, Y, Z = _grid(n)
    return (np.sqrt((X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2) - r).astype(np.float32)


def _torus(n, R=0.55, r=0.22):
    X, Y, Z = _grid(n)
    return (np.sqrt((np.sqrt(X ** 2 + Y ** 2) - R) ** 2 + Z ** 2) - r).astype(np.float32)


def _bumpy(n, seed=3):
    rng = np.random.default_rng(seed)
    X, Y, Z = _grid(n)
    f = np.sqrt(X ** 2 + Y ** 2 + Z ** 2) - 0.55
    for _ in range(6):
        k, ph = rng.uniform(2, 7, 3), rng.uniform(0, 6.28, 3)
        f = f + 0.05 * np.sin(k[0] * X + ph[0]) * np.sin(k[1] * Y + ph[1]) * np.sin(k[2] * Z + ph[2])
    return f.astype(np.float32)


def test_marching_tets_oracle_sphere_is_closed_and_outward():
    n = 24
    V, K = O.marching_tets_np(_sphere(n), 0.0)
    v, f = O.weld_triangles(V, K)
    m = TriMesh(v, f)
    h = 2.0 / (n - 1)
    assert m.is_watertight
    assert len(v) - 3 * len(f) // 2 + len(f) == 2                       # Euler characteristic of a sphere
    assert abs(m.area * h * h - 4 * np.pi * 0.36) < 0.02 * 4 * np.pi * 0.36
    ctr = (np.array([0.07, -0.03, 0.11]) + 1) / h
    nrm = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    assert ((nrm * (v[f].mean(1) - ctr)).sum(1) > 0).all()               # triangles face increasing field values
    r = np.linalg.norm(v * h - 1 - np.array([0.07, -0.03, 0.11]), axis=1)
    assert np.abs(r - 0.6).max() < 0.3 * h                               # linear interpolation error of a curved field


def test_marching_tets_oracle_torus_and_bumpy_are_closed():
    V, K = O.marching_tets_np(_torus(28), 0.0)
    v, f = O.weld_triangles(V, K)
    assert TriMesh(v, f).is_watertight and len(v) - 3 * len(f) // 2 + len(f) == 0      # genus 1
    V, K = O.marching_tets_np(_bumpy(20), 0.0)
    v, f = O.weld_triangles(V, K)
    assert TriMesh(v, f).is_watertight
    # a level that nothing reaches / a grid with values exactly on the level
    V, K = O.marching_tets_np(_sphere(8), -5.0)
    assert len(V) == 0
    g = np.round(_sphere(12) * 4) / 4
    V, K = O.marching_tets_np(g.astype(np.float32), 0.0)
    v, f = O.weld_triangles(V, K)
    assert TriMesh(v, f).is_watertight


def test_trimesh_container_roundtrip(tmp_path):
    V, K = O.marching_tets_np(_sphere(10), 0.0)
    v, f = O.weld_triangles(V, K)
    m = TriMesh(v, f)
    T = np.eye(4); T[:3, 3] = [1, 2, 3]; T[:3, :3] *= 0.5
    m2 = m.copy().apply_transform(T)
    np.testing.assert_allclose(m2.vertices, m.vertices * 0.5 + [1, 2, 3])
    assert abs(m2.area - 0.25 * m.area) < 1e-9
    for ext in ('obj', 'ply'):
        p = m.export(tmp_path / f'm.{ext}')
        txt = open(p).read()
        assert txt.count('\n') >= len(v) + len(f)
    with pytest.raises(ValueError):
        m.export(tmp_path / 'm.stl')


@pytest.mark.gpu
@pytest.mark.parametrize('name,n,iso', [('sphere', 24, 0.0), ('torus', 28, 0.0), ('bumpy', 33, 0.02), ('quantised', 12, 0.0)])
def test_marching_tets_kernel_matches_oracle(name, n, iso):
    from bundlesdf_b200 import _lib, ops
    f = {'sphere': _sphere, 'torus': _torus, 'bumpy': _bumpy, 'quantised': lambda n: (np.round(_sphere(n) * 4) / 4).astype(np.float32)}[name](n)
    if name == 'bumpy':
        f = f[:, :-3, :-7].copy()                                           # non-cubic grid
    V, K = O.marching_tets_np(f, iso)
    # raw kernel output (before welding): same triangles in the same order, bit-identical vertices
    lib = _lib.load()
    fd = torch.from_numpy(f).cuda()
    nx, ny, nz = f.shape
    counts = torch.empty((nx - 1) * (ny - 1) * (nz - 1), dtype=torch.int32, device='cuda')
    _lib.check(lib.nof_marching_tets_count(_lib.ptr(fd), nx, ny, nz, float(iso), _lib.ptr(counts), _lib.stream()), 'count')
    csum = torch.cumsum(counts, 0, dtype=torch.int64)
    assert int(csum[-1]) == len(V)
    offs = (csum - counts).contiguous()
    verts = torch.empty(len(V), 3, 3, device='cuda')
    keys = torch.empty(len(V), 3, dtype=torch.int64, device='cuda')
    _lib.check(lib.nof_marching_tets_emit(_lib.ptr(fd), nx, ny, nz, float(iso), _lib.ptr(offs), _lib.ptr(verts), _lib.ptr(keys), _lib.stream()), 'emit')
    np.testing.assert_array_equal(keys.cpu().numpy(), K)
    np.testing.assert_array_equal(verts.cpu().numpy(), V)
    # welded form through the public wrapper
    v, fa = ops.marching_tets(fd, iso)
    vo, fo = O.weld_triangles(V, K)
    np.testing.assert_array_equal(v.cpu().numpy(), vo)
    np.testing.assert_array_equal(fa.cpu().numpy(), fo)
    if name != 'bumpy':
        assert TriMesh(v.cpu().numpy(), fa.cpu().numpy()).is_watertight
