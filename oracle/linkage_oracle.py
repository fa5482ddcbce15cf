"""CPU restatement of the hierarchical clustering the reference calls through scipy.

TEST INFRASTRUCTURE ONLY (see gravity_oracle.py).  The reference's DistMat2Tree
(app/utils/dist_mat_to_tree.py:5-21) calls ``scipy.cluster.hierarchy.linkage(dist_list, method)``;
scipy is a third-party dependency of the reference (requirements.txt:4 pins scipy==1.14.0; this
image has 1.18.1), so its published algorithms are restated here, loop for loop, from
scipy/cluster/_hierarchy.pyx -- ``nn_chain`` (methods complete, average, weighted, ward),
``mst_single_linkage`` (single), ``fast_linkage`` + the binary ``Heap`` of _structures.pxi (centroid, median) and
``label`` -- and pinned bit for bit against the installed scipy by
tests/test_oracle_golden.py.  The device kernels of gravity2_b200/csrc/linkage.cu follow the same
loops; the comparison order is part of the result because real distance matrices are full of
exact ties.
"""
from __future__ import annotations

import math

import numpy as np


def _cidx(n: int, i: int, j: int) -> int:
    """condensed_index(n, i, j), scipy/cluster/_structures.pxi."""
    if i < j:
        return n * i - (i * (i + 1) // 2) + (j - i - 1)
    return n * j - (j * (j + 1) // 2) + (i - j - 1)


def new_dist(method: str, d_xi: float, d_yi: float, d_xy: float, nx: int, ny: int, ni: int) -> float:
    """linkage_methods[...] of scipy/cluster/_hierarchy_distance_update.pxi (each operation rounded on its own)."""
    if method == "average":
        return (nx * d_xi + ny * d_yi) / (nx + ny)
    if method == "complete":
        return max(d_xi, d_yi)
    if method == "weighted":
        return 0.5 * (d_xi + d_yi)
    if method == "ward":
        t = 1.0 / (nx + ny + ni)
        return math.sqrt((ni + nx) * t * d_xi * d_xi + (ni + ny) * t * d_yi * d_yi - ni * t * d_xy * d_xy)
    if method == "centroid":
        return _sqrt((((nx * d_xi * d_xi) + (ny * d_yi * d_yi)) - (nx * ny * d_xy * d_xy) / (nx + ny)) / (nx + ny))
    if method == "median":
        return _sqrt(0.5 * (d_xi * d_xi + d_yi * d_yi) - 0.25 * d_xy * d_xy)
    raise ValueError(method)


def _sqrt(v: float) -> float:
    """C sqrt: NaN for a negative argument (non-Euclidean distances reach it in centroid / median updates)."""
    return math.sqrt(v) if v >= 0 else math.nan


def label(Z: np.ndarray, n: int) -> None:
    """label() + LinkageUnionFind of scipy/cluster/_hierarchy.pyx: cluster ids of the sorted rows, sizes."""
    parent = list(range(2 * n - 1))
    size = [1] * (2 * n - 1)
    nxt = n

    def find(x):
        p = x
        while parent[x] != x:
            x = parent[x]
        while parent[p] != x:
            p, parent[p] = parent[p], x
        return x

    for i in range(n - 1):
        xr, yr = find(int(Z[i, 0])), find(int(Z[i, 1]))
        Z[i, 0], Z[i, 1] = (xr, yr) if xr < yr else (yr, xr)
        parent[xr] = parent[yr] = nxt
        size[nxt] = size[xr] + size[yr]
        Z[i, 3] = size[nxt]
        nxt += 1


def nn_chain(dists: np.ndarray, n: int, method: str) -> np.ndarray:
    """nn_chain of scipy/cluster/_hierarchy.pyx."""
    Z = np.empty((n - 1, 4))
    D = np.array(dists, dtype=np.float64)
    size = [1] * n
    chain = [0] * n
    chain_length = 0
    y = 0
    for k in range(n - 1):
        if chain_length == 0:
            chain_length = 1
            for i in range(n):
                if size[i] > 0:
                    chain[0] = i
                    break
        while True:
            x = chain[chain_length - 1]
            if chain_length > 1:
                y = chain[chain_length - 2]
                current_min = D[_cidx(n, x, y)]
            else:
                current_min = math.inf
            for i in range(n):
                if size[i] == 0 or x == i:
                    continue
                dist = D[_cidx(n, x, i)]
                if dist < current_min:
                    current_min = dist
                    y = i
            if chain_length > 1 and y == chain[chain_length - 2]:
                break
            chain[chain_length] = y
            chain_length += 1
        chain_length -= 2
        if x > y:
            x, y = y, x
        nx, ny = size[x], size[y]
        Z[k] = (x, y, current_min, nx + ny)
        size[x] = 0
        size[y] = nx + ny
        for i in range(n):
            ni = size[i]
            if ni == 0 or i == y:
                continue
            D[_cidx(n, i, y)] = new_dist(method, D[_cidx(n, i, x)], D[_cidx(n, i, y)], current_min, nx, ny, ni)
    Z = Z[np.argsort(Z[:, 2], kind="mergesort")]
    label(Z, n)
    return Z


def mst_single_linkage(dists: np.ndarray, n: int) -> np.ndarray:
    """mst_single_linkage of scipy/cluster/_hierarchy.pyx (method "single")."""
    Z = np.empty((n - 1, 4))
    merged = [0] * n
    D = [math.inf] * n
    x, y = 0, 0
    for k in range(n - 1):
        current_min = math.inf
        merged[x] = 1
        for i in range(n):
            if merged[i] == 1:
                continue
            dist = dists[_cidx(n, x, i)]
            if D[i] > dist:
                D[i] = dist
            if D[i] < current_min:
                y = i
                current_min = D[i]
        Z[k, 0], Z[k, 1], Z[k, 2] = x, y, current_min
        x = y
    Z = Z[np.argsort(Z[:, 2], kind="mergesort")]
    label(Z, n)
    return Z


class Heap:
    """Binary min-heap with key <-> position maps, scipy/cluster/_structures.pxi (the order of its swaps decides which of
    several equal minima comes out first, so it is part of the result)."""

    def __init__(self, values):
        self.size = len(values)
        self.index_by_key = list(range(self.size))
        self.key_by_index = list(range(self.size))
        self.values = list(values)
        for i in reversed(range(self.size // 2)):
            self.sift_down(i)

    def get_min(self):
        return self.key_by_index[0], self.values[0]

    def remove_min(self):
        self.swap(0, self.size - 1)
        self.size -= 1
        self.sift_down(0)

    def change_value(self, key, value):
        index = self.index_by_key[key]
        old_value = self.values[index]
        self.values[index] = value
        if value < old_value:
            self.sift_up(index)
        else:
            self.sift_down(index)

    def sift_up(self, index):
        parent = (index - 1) >> 1
        while index > 0 and self.values[parent] > self.values[index]:
            self.swap(index, parent)
            index = parent
            parent = (index - 1) >> 1

    def sift_down(self, index):
        child = (index << 1) + 1
        while child < self.size:
            if child + 1 < self.size and self.values[child + 1] < self.values[child]:
                child += 1
            if self.values[index] > self.values[child]:
                self.swap(index, child)
                index = child
                child = (index << 1) + 1
            else:
                break

    def swap(self, i, j):
        self.values[i], self.values[j] = self.values[j], self.values[i]
        key_i, key_j = self.key_by_index[i], self.key_by_index[j]
        self.key_by_index[i], self.key_by_index[j] = key_j, key_i
        self.index_by_key[key_i], self.index_by_key[key_j] = j, i


def _find_min_dist(n, D, size, x):
    """find_min_dist of scipy/cluster/_hierarchy.pyx: nearest live cluster among the higher-numbered ones."""
    current_min, y = math.inf, -1
    for i in range(x + 1, n):
        if size[i] == 0:
            continue
        dist = D[_cidx(n, x, i)]
        if dist < current_min:
            current_min, y = dist, i
    return y, current_min


def fast_linkage(dists: np.ndarray, n: int, method: str) -> np.ndarray:
    """fast_linkage of scipy/cluster/_hierarchy.pyx (methods centroid, median; Muellner's generic algorithm): per cluster a
    nearest-neighbour CANDIDATE among the higher-numbered clusters and a lower bound of its distance in a heap; the rows come
    out in merge order with cluster ids n + k (no sort, no relabelling: heights need not be monotone)."""
    Z = np.empty((n - 1, 4))
    D = np.array(dists, dtype=np.float64)
    size = [1] * n
    cluster_id = list(range(n))
    neighbor = [0] * (n - 1)
    min_dist = [0.0] * (n - 1)
    for x in range(n - 1):
        neighbor[x], min_dist[x] = _find_min_dist(n, D, size, x)
    heap = Heap(min_dist)
    x = y = 0
    dist = 0.0
    for k in range(n - 1):
        for _ in range(n - k):
            x, dist = heap.get_min()
            y = neighbor[x]
            if dist == D[_cidx(n, x, y)]:
                break
            y, dist = _find_min_dist(n, D, size, x)
            neighbor[x] = y
            min_dist[x] = dist
            heap.change_value(x, dist)
        heap.remove_min()
        id_x, id_y = cluster_id[x], cluster_id[y]
        nx, ny = size[x], size[y]
        if id_x > id_y:
            id_x, id_y = id_y, id_x
        Z[k] = (id_x, id_y, dist, nx + ny)
        size[x] = 0
        size[y] = nx + ny
        cluster_id[y] = n + k
        for z in range(n):
            nz = size[z]
            if nz == 0 or z == y:
                continue
            D[_cidx(n, z, y)] = new_dist(method, D[_cidx(n, z, x)], D[_cidx(n, z, y)], dist, nx, ny, nz)
        for z in range(x):
            if size[z] > 0 and neighbor[z] == x:
                neighbor[z] = y
        for z in range(y):
            if size[z] == 0:
                continue
            d = D[_cidx(n, z, y)]
            if d < min_dist[z]:
                neighbor[z] = y
                min_dist[z] = d
                heap.change_value(z, d)
        if y < n - 1:
            z, d = _find_min_dist(n, D, size, y)
            if z != -1:
                neighbor[y] = z
                min_dist[y] = d
                heap.change_value(y, d)
    return Z


def linkage(dists: np.ndarray, method: str = "average") -> np.ndarray:
    dists = np.asarray(dists, dtype=np.float64)
    n = int(round((1 + math.sqrt(1 + 8 * len(dists))) / 2))
    if method == "single":
        return mst_single_linkage(dists, n)
    if method in ("centroid", "median"):
        return fast_linkage(dists, n, method)
    return nn_chain(dists, n, method)
