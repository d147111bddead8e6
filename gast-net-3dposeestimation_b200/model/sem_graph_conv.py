"""Parameter shell for the non-channel-wise SemGraphConv of the reference
(`model/sem_graph_conv.py:10-55`; imported by nothing there, named by the north star).
It is the shared-`e` special case of SemCHGraphConv and runs on the same CUDA kernel
(shared-e flag + bias)."""
from __future__ import absolute_import, division

import math
import torch
import torch.nn as nn


class SemGraphConv(nn.Module):
    def __init__(self, in_features, out_features, adj, bias=True):
        super(SemGraphConv, self).__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.W = nn.Parameter(torch.zeros(size=(2, in_features, out_features), dtype=torch.float))
        nn.init.xavier_uniform_(self.W.data, gain=1.414)
        self.adj = adj
        self.m = (self.adj > 0)
        self.e = nn.Parameter(torch.zeros(1, int(self.m.sum().item()), dtype=torch.float))
        nn.init.constant_(self.e.data, 1)
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_features, dtype=torch.float))
            stdv = 1. / math.sqrt(self.W.size(2))
            self.bias.data.uniform_(-stdv, stdv)
        else:
            self.register_parameter('bias', None)

    def forward(self, input):
        from gast_b200 import engine
        return engine.run_semch(self, input)

    def __repr__(self):
        return self.__class__.__name__ + ' (' + str(self.in_features) + ' -> ' + str(self.out_features) + ')'
