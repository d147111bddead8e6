"""Skeleton -> dense adjacency (row a1 of the hot-path scope table).

Same contract as the reference `common/graph_utils.py:27-45`
(`adj_mx_from_skeleton(skeleton)` -> dense float32 (J,J) torch tensor: symmetrised
edges + self loops, row-normalised), built with dense numpy instead of scipy.sparse.
Only the `adj > 0` pattern (and the (J,J) shape) is consumed downstream.
"""
import numpy as np
import torch


def normalize(mx):
    """Row-normalise a dense matrix (rows that sum to 0 stay 0); cf. graph_utils.py:8-15."""
    mx = np.asarray(mx, dtype=np.float64)
    rowsum = mx.sum(1)
    with np.errstate(divide='ignore'):
        r_inv = np.where(rowsum != 0, 1.0 / rowsum, 0.0)
    return mx * r_inv[:, None]


def adj_mx_from_edges(num_pts, edges, sparse=False):
    """edges: iterable of (child, parent). cf. graph_utils.py:27-39 (dense result only)."""
    if sparse:
        raise NotImplementedError('the lifting path only uses the dense adjacency')
    a = np.zeros((num_pts, num_pts), dtype=np.float64)
    for i, j in edges:
        a[int(i), int(j)] += 1.0
    # symmetrise: keep max(a, a^T) entry-wise (graph_utils.py:33)
    a = np.maximum(a, a.T)
    a = normalize(a + np.eye(num_pts))
    return torch.tensor(a.astype(np.float32), dtype=torch.float)


def adj_mx_from_skeleton(skeleton):
    """cf. graph_utils.py:42-45."""
    n = skeleton.num_joints()
    edges = [(i, p) for i, p in zip(range(n), skeleton.parents()) if p >= 0]
    return adj_mx_from_edges(n, edges, sparse=False)
