"""Parameter shells for the global (non-local attention) half of a graph-attention block.

Drop-in for the reference `model/global_attention.py` (GlobalGraph :7-82, MultiGlobalGraph
:85-130, SingleGlobalGraph :133-173): identical parameter names, shapes and init.  Compute
is done by the sm_100a CUDA library through `gast_b200.engine`.
"""
from __future__ import absolute_import, division

import torch
from torch import nn


class GlobalGraph(nn.Module):
    """One non-local head: additive scores, LeakyReLU(0.2), softmax over all joints, learned
    bias C_k added after the softmax (global_attention.py:52-82)."""

    def __init__(self, adj, in_channels, inter_channels=None):
        super(GlobalGraph, self).__init__()
        self.adj = adj
        self.in_channels = in_channels
        self.inter_channels = inter_channels
        if self.inter_channels == self.in_channels // 2:
            self.g_channels = self.in_channels
        else:
            self.g_channels = self.inter_channels
        assert self.inter_channels > 0
        self.g = nn.Conv1d(self.in_channels, self.g_channels, kernel_size=1, stride=1, padding=0)
        self.theta = nn.Conv1d(self.in_channels, self.inter_channels, kernel_size=1, stride=1, padding=0)
        self.phi = nn.Conv1d(self.in_channels, self.inter_channels, kernel_size=1, stride=1, padding=0)
        self.C_k = nn.Parameter(torch.zeros(self.adj.shape, dtype=torch.float))
        self.concat_project = nn.Sequential(
            nn.Conv2d(self.inter_channels * 2, 1, 1, 1, 0, bias=False),
        )
        nn.init.kaiming_normal_(self.concat_project[0].weight)
        nn.init.kaiming_normal_(self.g.weight)
        nn.init.constant_(self.g.bias, 0)
        nn.init.kaiming_normal_(self.theta.weight)
        nn.init.constant_(self.theta.bias, 0)
        nn.init.kaiming_normal_(self.phi.weight)
        nn.init.constant_(self.phi.bias, 0)

    def forward(self, x):
        # x: (B*T, C, N) like the reference; returns (B*T, g_channels, N)
        from gast_b200 import engine
        return engine.run_global_head(self, x)


class MultiGlobalGraph(nn.Module):
    """in//inter heads -> cat -> 1x1 conv -> BN -> ReLU -> Dropout (global_attention.py:85-130)."""

    def __init__(self, adj, in_channels, inter_channels, dropout=None):
        super(MultiGlobalGraph, self).__init__()
        self.num_non_local = in_channels // inter_channels
        self.attentions = nn.ModuleList(
            [GlobalGraph(adj, in_channels, inter_channels) for _ in range(self.num_non_local)])
        self.cat_conv = nn.Conv2d(in_channels, in_channels, 1, bias=False)
        self.cat_bn = nn.BatchNorm2d(in_channels, momentum=0.1)
        self.relu = nn.ReLU(inplace=True)
        self.dropout = nn.Dropout(dropout) if dropout is not None else None

    def forward(self, x):
        from gast_b200 import engine
        return engine.run_multi_global(self, x)


class SingleGlobalGraph(nn.Module):
    """Kept for import compatibility only: the reference never instantiates it (it is named
    in a comment at gast_net.py:17).  Parameters match global_attention.py:133-173; calling
    it raises, because no kernel is built for dead code."""

    def __init__(self, adj, in_channels, output_channels, dropout=None):
        super(SingleGlobalGraph, self).__init__()
        self.attentions = GlobalGraph(adj, in_channels, output_channels // 2)
        self.bn = nn.BatchNorm2d(in_channels, momentum=0.1)
        self.relu = nn.ReLU(inplace=True)
        self.dropout = nn.Dropout(dropout) if dropout is not None else None

    def forward(self, x):
        raise NotImplementedError('SingleGlobalGraph is dead code in the reference; not on the lifting path')
