"""Skeleton -> adjacency (row a1 of the hot-path scope table).

Same contract as the reference `common/graph_utils.py` -- this module shadows it for every unchanged caller
(`common` is a namespace package on both sides), so all of its public functions keep their signatures, defaults
and results: `adj_mx_from_skeleton(skeleton)` -> dense float32 (J,J) torch tensor (symmetrised edges + self
loops, row-normalised, graph_utils.py:42-45), `adj_mx_from_edges(num_pts, edges, sparse=True)` -> torch sparse COO
tensor by default like graph_utils.py:27-39, `normalize`, `sparse_mx_to_torch_sparse_tensor`.  Built with dense
numpy in float32 (the reference's scipy.sparse arithmetic is float32 too) -- the lifting path only consumes the
`adj > 0` pattern and the (J,J) shape.
"""
import numpy as np
import torch


def normalize(mx):
    """Row-normalise a matrix (rows that sum to 0 stay 0); cf. graph_utils.py:8-15.  Accepts a dense array or a
    scipy.sparse matrix and returns the same kind."""
    if hasattr(mx, 'tocoo'):                                   # scipy.sparse input, as the reference passes
        import scipy.sparse as sp
        rowsum = np.array(mx.sum(1))
        with np.errstate(divide='ignore'):
            r_inv = np.power(rowsum, -1).flatten()
        r_inv[np.isinf(r_inv)] = 0.
        return sp.diags(r_inv).dot(mx)
    mx = np.asarray(mx)
    rowsum = mx.sum(1)
    with np.errstate(divide='ignore'):
        r_inv = np.power(rowsum, -1)
    r_inv[np.isinf(r_inv)] = 0.
    return r_inv[:, None] * mx


def sparse_mx_to_torch_sparse_tensor(sparse_mx):
    """scipy.sparse (or dense) matrix -> torch sparse COO float tensor, row-major entry order; cf. graph_utils.py:18-24."""
    if hasattr(sparse_mx, 'tocoo'):
        coo = sparse_mx.tocoo().astype(np.float32)
        row, col, data = coo.row, coo.col, coo.data
        shape = coo.shape
    else:
        dense = np.asarray(sparse_mx, dtype=np.float32)
        row, col = np.nonzero(dense)
        data, shape = dense[row, col], dense.shape
    indices = torch.from_numpy(np.vstack((row, col)).astype(np.int64))
    return torch.sparse_coo_tensor(indices, torch.from_numpy(np.ascontiguousarray(data)), torch.Size(shape))


def _dense_adjacency(num_pts, edges):
    edges = np.array(list(edges), dtype=np.int32).reshape(-1, 2)
    a = np.zeros((num_pts, num_pts), dtype=np.float32)
    np.add.at(a, (edges[:, 0], edges[:, 1]), np.float32(1.0))  # duplicate edges add up, like coo_matrix
    # symmetrise: entry-wise max(a, a^T) (graph_utils.py:33)
    a = np.maximum(a, a.T)
    return normalize(a + np.eye(num_pts, dtype=np.float32)).astype(np.float32)


def adj_mx_from_edges(num_pts, edges, sparse=True):
    """edges: iterable of (child, parent); cf. graph_utils.py:27-39 (same default: sparse=True)."""
    a = _dense_adjacency(num_pts, edges)
    if sparse:
        return sparse_mx_to_torch_sparse_tensor(a)
    return torch.tensor(a, dtype=torch.float)


def adj_mx_from_skeleton(skeleton):
    """cf. graph_utils.py:42-45."""
    n = skeleton.num_joints()
    edges = [(i, p) for i, p in zip(range(n), skeleton.parents()) if p >= 0]
    return adj_mx_from_edges(n, edges, sparse=False)
