"""B200-native drop-in for the reference `model/gast_net.py`.

`from model.gast_net import *` must keep working for the reference callers (`main.py:7`,
`reconstruction.py:15`, `gen_skes.py:13`), including the fact that `torch` and `nn` reach
`trainval.py` through this star import (`trainval.py:60` uses `nn.DataParallel`).  So this
module exposes `torch`, `nn`, `GraphAttentionBlock`, `SpatioTemporalModelBase`,
`SpatioTemporalModel`, `SpatioTemporalModelOptimized1f` (+ the attention classes), and does
NOT import `torchsummary`.

The classes are thin `nn.Module` shells: they own ordinary parameters/buffers under the
reference's exact names (so `state_dict`, optimisers and checkpoints are interchangeable,
gast_net.py:113-157,191-234) and `forward()` enqueues hand-written sm_100a kernels through
the C-ABI library (`include/gast_b200.h`).  There is no torch-op or CPU fallback: a missing
library or a CPU tensor raises.
"""
import torch
import torch.nn as nn
from model.local_attention import LocalGraph
from model.global_attention import MultiGlobalGraph, SingleGlobalGraph


class GraphAttentionBlock(nn.Module):
    """cat[x, Local(x), Global(x)] -> 1x1 conv 3C->2C -> BN -> ReLU (gast_net.py:8-33)."""

    def __init__(self, adj, input_dim, output_dim, p_dropout):
        super(GraphAttentionBlock, self).__init__()
        hid_dim = output_dim
        self.relu = nn.ReLU(inplace=True)
        self.local_graph_layer = LocalGraph(adj, input_dim, hid_dim, p_dropout)
        self.global_graph_layer = MultiGlobalGraph(adj, input_dim, input_dim // 4, dropout=p_dropout)
        self.cat_conv = nn.Conv2d(3 * output_dim, 2 * output_dim, 1, bias=False)
        self.cat_bn = nn.BatchNorm2d(2 * output_dim, momentum=0.1)

    def forward(self, x):
        # x: (B, C, T, N) -> (B, 2C, T, N), like the reference
        from gast_b200 import engine
        return engine.run_block(self, x)


class SpatioTemporalModelBase(nn.Module):
    """Do not instantiate this class (gast_net.py:36-104)."""

    def __init__(self, adj, num_joints_in, in_features, num_joints_out,
                 filter_widths, causal, dropout, channels):
        super().__init__()
        for fw in filter_widths:
            assert fw % 2 != 0, 'Only odd filter widths are supported'
        self.num_joints_in = num_joints_in
        self.in_features = in_features
        self.num_joints_out = num_joints_out
        self.filter_widths = filter_widths
        self.drop = nn.Dropout(dropout)
        self.relu = nn.ReLU(inplace=True)
        self.pad = [filter_widths[0] // 2]
        self.init_bn = nn.BatchNorm2d(in_features, momentum=0.1)
        self.expand_bn = nn.BatchNorm2d(channels, momentum=0.1)
        self.shrink = nn.Conv2d(2 ** len(self.filter_widths) * channels, 3, 1, bias=False)
        # engine-side description of this instance (not part of state_dict)
        self._gast_channels = channels
        self._gast_causal = bool(causal)
        self._gast_dropout = float(dropout)
        self._gast_adj = adj

    def receptive_field(self):
        """Total receptive field in frames (gast_net.py:62-69)."""
        return 1 + 2 * sum(self.pad)

    def total_causal_shift(self):
        """Asymmetric padding offset; kept bug-for-bug with gast_net.py:71-82."""
        frames = self.causal_shift[0]
        next_dilation = self.filter_widths[0]
        for i in range(1, len(self.filter_widths)):
            frames += self.causal_shift[i] * next_dilation
            next_dilation *= self.filter_widths[i]
        return frames

    def _build_layers(self, adj, filter_widths, channels, dropout, causal, strided, dense):
        layers_conv, layers_bn = [], []
        layers_graph_conv = [GraphAttentionBlock(adj, channels, channels, p_dropout=dropout)]
        self.causal_shift = [(filter_widths[0] // 2) if causal else 0]
        next_dilation = filter_widths[0]
        for i in range(1, len(filter_widths)):
            width = 2 ** i * channels
            self.pad.append((filter_widths[i] - 1) * next_dilation // 2)
            if strided:
                self.causal_shift.append((filter_widths[i] // 2) if causal else 0)
                layers_conv.append(nn.Conv2d(width, width, (filter_widths[i], 1),
                                             stride=(filter_widths[i], 1), bias=False))
            else:
                self.causal_shift.append((filter_widths[i] // 2 * next_dilation) if causal else 0)
                layers_conv.append(nn.Conv2d(width, width,
                                             (filter_widths[i], 1) if not dense else (2 * self.pad[-1] + 1, 1),
                                             dilation=(next_dilation, 1) if not dense else (1, 1), bias=False))
            layers_bn.append(nn.BatchNorm2d(width, momentum=0.1))
            layers_conv.append(nn.Conv2d(width, width, 1, dilation=1, bias=False))
            layers_bn.append(nn.BatchNorm2d(width, momentum=0.1))
            layers_graph_conv.append(GraphAttentionBlock(adj, width, width, p_dropout=dropout))
            next_dilation *= filter_widths[i]
        self.layers_conv = nn.ModuleList(layers_conv)
        self.layers_bn = nn.ModuleList(layers_bn)
        self.layers_graph_conv = nn.ModuleList(layers_graph_conv)

    def forward(self, x):
        """x: (B, T, N, C=in_features) float32 -> (B, T_out, N, 3) (gast_net.py:84-104)."""
        assert len(x.shape) == 4
        assert x.shape[-2] == self.num_joints_in
        assert x.shape[-1] == self.in_features
        from gast_b200 import engine
        return engine.run_model(self, x)


class SpatioTemporalModel(SpatioTemporalModelBase):
    """General model: dilated temporal convolutions, any sequence length >= receptive field
    (gast_net.py:107-177)."""

    def __init__(self, adj, num_joints_in, in_features, num_joints_out,
                 filter_widths, causal=False, dropout=0.25, channels=64, dense=False):
        super().__init__(adj, num_joints_in, in_features, num_joints_out, filter_widths, causal, dropout, channels)
        self.expand_conv = nn.Conv2d(in_features, channels, (filter_widths[0], 1), bias=False)
        nn.init.kaiming_normal_(self.expand_conv.weight)
        self._gast_strided = False
        self._gast_dense = bool(dense)
        self._build_layers(adj, filter_widths, channels, dropout, causal, strided=False, dense=dense)


class SpatioTemporalModelOptimized1f(SpatioTemporalModelBase):
    """Single-output-frame model: strided instead of dilated convolutions; weights are
    interchangeable with SpatioTemporalModel (gast_net.py:180-251)."""

    def __init__(self, adj, num_joints_in, in_features, num_joints_out,
                 filter_widths, causal=False, dropout=0.25, channels=64):
        super().__init__(adj, num_joints_in, in_features, num_joints_out, filter_widths, causal, dropout, channels)
        self.expand_conv = nn.Conv2d(in_features, channels, (filter_widths[0], 1),
                                     stride=(filter_widths[0], 1), bias=False)
        nn.init.kaiming_normal_(self.expand_conv.weight)
        self._gast_strided = True
        self._gast_dense = False
        self._build_layers(adj, filter_widths, channels, dropout, causal, strided=True, dense=False)
