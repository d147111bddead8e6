"""Parameter shells for the local (semantic graph-conv) half of a graph-attention block.

Drop-in for the reference `model/local_attention.py`: same class names, constructor
arguments, parameter names/shapes (`W (2,Cin,Cout)`, `e (Cout,nnz)`, `bn_1`, `bn_2`,
`cat_conv`, `cat_bn`) and init (local_attention.py:20-26,116-123), so `state_dict`s are
interchangeable.  The arithmetic is NOT here: `forward` hands the tensors to the sm_100a
CUDA library through `gast_b200.engine` (no torch-op fallback).
"""
from __future__ import absolute_import, division

import math
import numpy as np
import torch
import torch.nn as nn

_JOINT_TABLES = {
    # J: (distal, left, right)  -- values from the reference tables, local_attention.py:65-87
    17: ([3, 6, 10, 13, 16], [4, 5, 6, 11, 12, 13], [1, 2, 3, 14, 15, 16]),
    16: ([3, 6, 9, 12, 15], [4, 5, 6, 10, 11, 12], [1, 2, 3, 13, 14, 15]),
    15: ([4, 7, 10, 13], [2, 3, 4, 8, 9, 10], [5, 6, 7, 11, 12, 13]),
    19: ([3, 4, 7, 8, 12, 15, 18], [5, 6, 7, 8, 13, 14, 15], [1, 2, 3, 4, 16, 17, 18]),
}


def local_adjacencies(adj):
    """(adj_sym, adj_con) float tensors as LocalGraph.__init__ builds them
    (local_attention.py:92-114): sym = I + left<->right partner; con = rows of adj for
    non-distal joints, rows of adj@adj for distal joints."""
    J = adj.shape[0]
    if J not in _JOINT_TABLES:
        raise KeyError("The dimension of adj matrix is wrong!")
    distal, left, right = _JOINT_TABLES[J]
    a = adj.detach().cpu().to(torch.float32)
    sym = torch.eye(J, dtype=torch.float32)
    for l, r in zip(left, right):
        sym[l, r] = 1.0
        sym[r, l] = 1.0
    first = a.clone()
    second = a @ a
    is_distal = torch.zeros(J, dtype=torch.bool)
    is_distal[distal] = True
    con = torch.where(is_distal[:, None], second, first)
    return sym, con


class SemCHGraphConv(nn.Module):
    """Semantic channel-wise graph convolution (reference local_attention.py:10-56)."""

    def __init__(self, in_features, out_features, adj, bias=False):
        super(SemCHGraphConv, self).__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.W = nn.Parameter(torch.zeros(size=(2, in_features, out_features), dtype=torch.float))
        nn.init.xavier_uniform_(self.W.data, gain=1.414)
        # plain attributes (not buffers), like the reference (local_attention.py:23-24)
        self.adj = adj
        self.m = (adj > 0)
        self.e = nn.Parameter(torch.zeros(out_features, int(self.m.sum().item()), dtype=torch.float))
        nn.init.constant_(self.e.data, 1)
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_features, dtype=torch.float))
            stdv = 1. / math.sqrt(self.W.size(1))
            self.bias.data.uniform_(-stdv, stdv)
        else:
            self.register_parameter('bias', None)

    def forward(self, input):
        from gast_b200 import engine
        return engine.run_semch(self, input)

    def __repr__(self):
        return self.__class__.__name__ + ' (' + str(self.in_features) + ' -> ' + str(self.out_features) + ')'


class LocalGraph(nn.Module):
    """Two SemCH graph convs (symmetry / connectivity) -> BN/ReLU -> cat -> 1x1 -> BN -> ReLU
    -> Dropout (reference local_attention.py:59-151)."""

    def __init__(self, adj, input_dim, output_dim, dropout=None):
        super(LocalGraph, self).__init__()
        adj_sym, adj_con = local_adjacencies(adj)
        self.gcn_sym = SemCHGraphConv(input_dim, output_dim, adj_sym)
        self.bn_1 = nn.BatchNorm2d(output_dim, momentum=0.1)
        self.gcn_con = SemCHGraphConv(input_dim, output_dim, adj_con)
        self.bn_2 = nn.BatchNorm2d(output_dim, momentum=0.1)
        self.relu = nn.ReLU()
        self.cat_conv = nn.Conv2d(2 * output_dim, output_dim, 1, bias=False)
        self.cat_bn = nn.BatchNorm2d(output_dim, momentum=0.1)
        self.dropout = nn.Dropout(dropout) if dropout is not None else None

    def forward(self, input):
        from gast_b200 import engine
        return engine.run_local(self, input)
