"""Minimal triangle-mesh container returned by NerfRunner.extract_mesh when `trimesh` (what the reference returns,
nerf_runner.py:1404) is not installed: the attributes and methods the reference's callers touch (bundlesdf.py:234-240, 747-763:
.vertices, .faces, .apply_transform, .export, .copy) on plain numpy arrays."""
import numpy as np


class TriMesh:
    def __init__(self, vertices, faces, vertex_colors=None):
        self.vertices = np.asarray(vertices, dtype=np.float64).reshape(-1, 3)
        self.faces = np.asarray(faces, dtype=np.int64).reshape(-1, 3)
        self.vertex_colors = None if vertex_colors is None else np.asarray(vertex_colors)

    def copy(self):
        return TriMesh(self.vertices.copy(), self.faces.copy(), None if self.vertex_colors is None else self.vertex_colors.copy())

    def apply_transform(self, T):
        T = np.asarray(T, dtype=np.float64)
        self.vertices = self.vertices @ T[:3, :3].T + T[:3, 3]
        return self

    def merge_vertices(self):
        """trimesh.Trimesh.merge_vertices (bundlesdf.py:748): weld vertices with identical positions, re-index the faces, drop
        the faces that become degenerate. extract_mesh() already welds by grid-edge key, so this is normally a no-op."""
        uniq, inv = np.unique(self.vertices, axis=0, return_inverse=True)
        inv = inv.reshape(-1)
        if len(uniq) == len(self.vertices):
            return self
        if self.vertex_colors is not None:
            first = np.full(len(uniq), -1, dtype=np.int64)
            first[inv[::-1]] = np.arange(len(inv))[::-1]
            self.vertex_colors = self.vertex_colors[first]
        f = inv[self.faces]
        self.vertices = uniq
        self.faces = f[(f[:, 0] != f[:, 1]) & (f[:, 1] != f[:, 2]) & (f[:, 0] != f[:, 2])]
        return self

    @property
    def edges(self):
        """[3F,2] directed edges, trimesh order (what Utils.trimesh_split feeds to connected_components, Utils.py:290)."""
        f = self.faces
        return np.stack([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 1).reshape(-1, 2)

    def update_vertices(self, mask):
        """trimesh.Trimesh.update_vertices(mask) (Utils.py:296): keep the masked vertices and the faces that only use them."""
        mask = np.asarray(mask, dtype=bool)
        remap = np.cumsum(mask) - 1
        keep = mask[self.faces].all(axis=1)
        self.faces = remap[self.faces[keep]]
        self.vertices = self.vertices[mask]
        if self.vertex_colors is not None:
            self.vertex_colors = self.vertex_colors[mask]
        return self

    def split(self, min_edge=0):
        """Connected components as separate meshes (role of Utils.trimesh_split, Utils.py:287-298: components with fewer than
        `min_edge` vertices are dropped) — union-find over the edges, no trimesh/networkx needed."""
        n = len(self.vertices)
        parent = np.arange(n)
        e = self.edges
        for _ in range(64):                                  # pointer-jumping label propagation; converges in O(log diameter)
            lo = np.minimum(parent[e[:, 0]], parent[e[:, 1]])
            new = parent.copy()
            np.minimum.at(new, e[:, 0], lo)
            np.minimum.at(new, e[:, 1], lo)
            new = new[new]
            if np.array_equal(new, parent):
                break
            parent = new
        out = []
        used = np.zeros(n, dtype=bool)
        used[self.faces.reshape(-1)] = True
        for lab in np.unique(parent[used]):
            mask = (parent == lab) & used
            if mask.sum() < max(min_edge, 1):
                continue
            out.append(self.copy().update_vertices(mask))
        return out

    @property
    def area(self):
        a, b, c = (self.vertices[self.faces[:, i]] for i in range(3))
        return float(0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1).sum())

    @property
    def is_watertight(self):
        """Every directed edge occurs once and its reverse occurs once (closed, consistently oriented 2-manifold)."""
        f = self.faces
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0)
        n = int(self.vertices.shape[0]) + 1
        fwd = e[:, 0] * n + e[:, 1]
        rev = e[:, 1] * n + e[:, 0]
        u, cnt = np.unique(fwd, return_counts=True)
        return bool(len(f) > 0 and (cnt == 1).all() and np.array_equal(u, np.unique(rev)))

    def export(self, path):
        path = str(path)
        if path.endswith('.obj'):
            with open(path, 'w') as fh:
                for i, v in enumerate(self.vertices):
                    if self.vertex_colors is not None:
                        c = np.asarray(self.vertex_colors[i][:3], dtype=np.float64) / (255.0 if self.vertex_colors.dtype == np.uint8 else 1.0)
                        fh.write(f'v {v[0]:.8f} {v[1]:.8f} {v[2]:.8f} {c[0]:.6f} {c[1]:.6f} {c[2]:.6f}\n')
                    else:
                        fh.write(f'v {v[0]:.8f} {v[1]:.8f} {v[2]:.8f}\n')
                for f in self.faces + 1:
                    fh.write(f'f {f[0]} {f[1]} {f[2]}\n')
        elif path.endswith('.ply'):
            with open(path, 'w') as fh:
                fh.write('ply\nformat ascii 1.0\n')
                fh.write(f'element vertex {len(self.vertices)}\nproperty float x\nproperty float y\nproperty float z\n')
                fh.write(f'element face {len(self.faces)}\nproperty list uchar int vertex_indices\nend_header\n')
                for v in self.vertices:
                    fh.write(f'{v[0]:.8f} {v[1]:.8f} {v[2]:.8f}\n')
                for f in self.faces:
                    fh.write(f'3 {f[0]} {f[1]} {f[2]}\n')
        else:
            raise ValueError(f'TriMesh.export: unsupported extension in {path!r} (.obj, .ply)')
        return path
