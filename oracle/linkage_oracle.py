"""CPU restatement of the hierarchical clustering the reference calls through scipy.

TEST INFRASTRUCTURE ONLY (see gravity_oracle.py).  The reference's DistMat2Tree
(app/utils/dist_mat_to_tree.py:5-21) calls ``scipy.cluster.hierarchy.linkage(dist_list, method)``;
scipy is a third-party dependency of the reference (requirements.txt:4 pins scipy==1.14.0; this
image has 1.18.1), so its published algorithms are restated here, loop for loop, from
scipy/cluster/_hierarchy.pyx -- ``nn_chain`` (methods complete, average, weighted, ward),
``mst_single_linkage`` (single) and ``label`` -- and pinned bit for bit against the installed scipy by
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
    raise ValueError(method)


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


def linkage(dists: np.ndarray, method: str = "average") -> np.ndarray:
    dists = np.asarray(dists, dtype=np.float64)
    n = int(round((1 + math.sqrt(1 + 8 * len(dists))) / 2))
    return mst_single_linkage(dists, n) if method == "single" else nn_chain(dists, n, method)
