"""numpy restatement of the reference's MISE octree (utils/libmise/mise.pyx) and of the vertex rule of its marching
cubes (utils/libmcubes/marchingcubes.h), in the dense-array formulation the CUDA kernels use.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned against the reference's own compiled modules (oracle/_ref, built
by oracle/build_ref.py) in tests/test_oracle.py.

  MiseNP(resolution_0, depth, threshold): query() -> int64 [n,3]; update(points, values); to_dense() -> float64 grid
      mise.pyx:42-84 (initial voxels / grid points), :86-103 update, :105-127 query, :129-168 to_dense,
      :189-242 subdivide_voxels (a leaf voxel below the maximum depth is subdivided when the KNOWN grid points on its
      closed cube include a value >= threshold and a value <= threshold), :244-286 subdivide_voxel.
  mc_vertices(volume, iso): the set of marching-cubes vertices -- one per grid edge whose ends differ in
      ``value <= iso`` (marchingcubes.h:65-68), at the linearly interpolated position, midpoint when the ends are equal
      (marchingcubes.cpp:290-297), shifted by +0.5 like libmcubes does (marchingcubes.h:45 ``+ dx/2``).
"""
from __future__ import annotations

import numpy as np


class MiseNP:
    def __init__(self, resolution_0: int, depth: int, threshold: float):
        self.resolution_0, self.depth, self.threshold = resolution_0, depth, threshold
        self.resolution = R = resolution_0 << depth
        self.state = np.zeros((R + 1,) * 3, np.uint8)      # 0 absent, 1 unknown, 2 known
        self.val = np.zeros((R + 1,) * 3, np.float64)
        self.level = np.zeros((R,) * 3, np.uint8)          # level of the leaf voxel containing each unit cell
        step = 1 << depth
        self.state[::step, ::step, ::step] = 1

    def query(self) -> np.ndarray:
        return np.argwhere(self.state == 1).astype(np.int64)

    def update(self, points: np.ndarray, values: np.ndarray) -> None:
        p = np.asarray(points, np.int64)
        self.val[p[:, 0], p[:, 1], p[:, 2]] = values
        self.state[p[:, 0], p[:, 1], p[:, 2]] = 2
        self._subdivide()

    def _subdivide(self) -> None:
        R, depth, thr = self.resolution, self.depth, self.threshold
        pos = np.zeros((R,) * 3, bool)
        neg = np.zeros((R,) * 3, bool)
        known = np.argwhere(self.state == 2)
        v = self.val[known[:, 0], known[:, 1], known[:, 2]]
        for a in (-1, 0):
            for b in (-1, 0):
                for c in (-1, 0):
                    cell = known + np.array([a, b, c])
                    ok = ((cell >= 0) & (cell < R)).all(1)
                    cc, vv = cell[ok], v[ok]
                    sh = depth - self.level[cc[:, 0], cc[:, 1], cc[:, 2]].astype(np.int64)
                    o = (cc >> sh[:, None]) << sh[:, None]          # origin cell of the leaf voxel
                    pos[o[vv >= thr, 0], o[vv >= thr, 1], o[vv >= thr, 2]] = True
                    neg[o[vv <= thr, 0], o[vv <= thr, 1], o[vv <= thr, 2]] = True
        cells = np.argwhere(self.level < depth)
        lv = self.level[cells[:, 0], cells[:, 1], cells[:, 2]].astype(np.int64)
        sh = depth - lv
        o = (cells >> sh[:, None]) << sh[:, None]
        act = pos[o[:, 0], o[:, 1], o[:, 2]] & neg[o[:, 0], o[:, 1], o[:, 2]]
        cells, o, sh, lv = cells[act], o[act], sh[act], lv[act]
        is_origin = (cells == o).all(1)
        oo, h = o[is_origin], (1 << sh[is_origin]) >> 1
        for a in range(3):
            for b in range(3):
                for c in range(3):
                    p = oo + np.stack([a * h, b * h, c * h], 1)
                    new = self.state[p[:, 0], p[:, 1], p[:, 2]] == 0
                    self.state[p[new, 0], p[new, 1], p[new, 2]] = 1
        self.level[cells[:, 0], cells[:, 1], cells[:, 2]] = lv + 1

    def to_dense(self) -> np.ndarray:
        out = np.where(self.state == 2, self.val, np.nan)
        n = self.resolution + 1
        for axis in range(3):
            for i in range(1, n):
                cur = np.take(out, i, axis)
                prev = np.take(out, i - 1, axis)
                fill = np.isnan(cur)
                cur = np.where(fill, prev, cur)
                idx = [slice(None)] * 3
                idx[axis] = i
                out[tuple(idx)] = cur
        return out


def mc_vertices(volume: np.ndarray, iso: float) -> np.ndarray:
    """Sorted [V,3] float64 vertex positions (libmcubes frame: grid index + 0.5)."""
    vol = np.asarray(volume, np.float64)
    s = vol <= iso
    out = []
    for axis in range(3):
        a = [slice(None)] * 3
        b = [slice(None)] * 3
        a[axis], b[axis] = slice(0, -1), slice(1, None)
        f0, f1 = vol[tuple(a)], vol[tuple(b)]
        cross = s[tuple(a)] != s[tuple(b)]
        idx = np.argwhere(cross).astype(np.float64)
        g0, g1 = f0[cross], f1[cross]
        with np.errstate(divide="ignore", invalid="ignore"):
            t = np.where(g1 == g0, 0.5, (iso - g0) / (g1 - g0))
        idx[:, axis] += t
        out.append(idx + 0.5)
    v = np.concatenate(out, 0)
    return v[np.lexsort((v[:, 2], v[:, 1], v[:, 0]))]
