"""ORACLE (test infrastructure, not product): numpy restatement of the GAST-Net lifting path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline leg may import this
file.  The product path (gast-net-3dposeestimation_b200/) never does.

Every function restates, op for op, what the reference executes on CPU in fp32 and cites
the reference file:line it follows.  The restatement is deliberately literal (it
materialises the concatenated pairwise feature of the global attention, the per-channel
(C,J,J) adjacency of the channel-wise graph conv, ...) so that (i) it is an independent
check of the algebraic shortcuts the CUDA path takes and (ii) its CPU time is representative
of the reference's own implementation.

PINNING: the reference holds no tests or golden vectors for this path, so this
oracle is pinned against outputs of the reference itself, generated in the build container
by importing /root/reference (tests/golden/make_golden.py -> tests/golden/*.npz) and checked
by tests/test_oracle_golden.py.

Parameter convention: `p` is a dict {state_dict key: numpy array} exactly as
`module.state_dict()` names them; `prefix` selects a sub-module.
"""
import numpy as np

BN_EPS = 1e-5  # torch.nn.BatchNorm2d default, used everywhere in the reference


# --------------------------------------------------------------------------------------
# graph construction (host side, integer/boolean)
# --------------------------------------------------------------------------------------
def adj_from_parents(parents):
    """common/graph_utils.py:27-45: edges (i,parent) -> symmetrise -> +I -> row-normalise."""
    J = len(parents)
    a = np.zeros((J, J), dtype=np.float64)
    for i, p in enumerate(parents):
        if p >= 0:
            a[i, p] = 1.0
    a = np.maximum(a, a.T)
    a = a + np.eye(J)
    a = a / a.sum(1, keepdims=True)
    return a.astype(np.float32)


_TABLES = {  # model/local_attention.py:65-87
    17: ([3, 6, 10, 13, 16], [4, 5, 6, 11, 12, 13], [1, 2, 3, 14, 15, 16]),
    16: ([3, 6, 9, 12, 15], [4, 5, 6, 10, 11, 12], [1, 2, 3, 13, 14, 15]),
    15: ([4, 7, 10, 13], [2, 3, 4, 8, 9, 10], [5, 6, 7, 11, 12, 13]),
    19: ([3, 4, 7, 8, 12, 15, 18], [5, 6, 7, 8, 13, 14, 15], [1, 2, 3, 4, 16, 17, 18]),
}


def local_masks(adj):
    """model/local_attention.py:92-114 -> boolean (J,J) masks (sym, con)."""
    J = adj.shape[0]
    if J not in _TABLES:
        raise KeyError("The dimension of adj matrix is wrong!")
    distal, left, right = _TABLES[J]
    sym = np.zeros((J, J), dtype=np.float32)
    for i in range(J):
        sym[i, i] = 1
        if i in left:
            sym[i, right[left.index(i)]] = 1
        if i in right:
            sym[i, left[right.index(i)]] = 1
    first = adj.copy()
    second = adj @ adj
    for i in range(J):
        if i in distal:
            first[i] = 0
        else:
            second[i] = 0
    con = first + second
    return sym > 0, con > 0


# --------------------------------------------------------------------------------------
# elementary ops, fp32
# --------------------------------------------------------------------------------------
def _softmax(x, axis):
    m = x.max(axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=axis, keepdims=True)


def batchnorm(x, p, prefix, training=False, caxis=1):
    """torch.nn.BatchNorm2d on channel axis `caxis`.  eval: running stats; train: batch
    mean / biased variance over all other axes.  Returns (y, stats) where stats is
    (mean, biased_var, count) in training mode, else None."""
    shp = [1] * x.ndim
    shp[caxis] = -1
    w = p[prefix + 'weight'].reshape(shp)
    b = p[prefix + 'bias'].reshape(shp)
    if training:
        axes = tuple(a for a in range(x.ndim) if a != caxis)
        mean = x.mean(axis=axes, dtype=np.float64).astype(np.float32)
        var = x.var(axis=axes, dtype=np.float64).astype(np.float32)
        n = x.size // x.shape[caxis]
        stats = (mean, var, n)
    else:
        mean = p[prefix + 'running_mean']
        var = p[prefix + 'running_var']
        stats = None
    y = (x - mean.reshape(shp)) / np.sqrt(var.reshape(shp) + np.float32(BN_EPS)) * w + b
    return y.astype(np.float32), stats


def relu(x):
    return np.maximum(x, 0)


def conv1x1(x, w):
    """nn.Conv2d(Cin,Cout,1,bias=False) on (B,C,T,N); w: (Cout,Cin,1,1)."""
    w2 = w.reshape(w.shape[0], w.shape[1])
    return np.einsum('oc,bctn->botn', w2, x, optimize=True).astype(np.float32)


def conv_t(x, w, stride=1, dilation=1):
    """nn.Conv2d(Cin,Cout,(k,1),stride=(stride,1),dilation=(dilation,1),bias=False), no padding."""
    Co, Ci, k, _ = w.shape
    B, C, T, N = x.shape
    span = (k - 1) * dilation + 1
    T_out = (T - span) // stride + 1
    y = np.zeros((B, Co, T_out, N), dtype=np.float32)
    for kk in range(k):
        xs = x[:, :, kk * dilation: kk * dilation + (T_out - 1) * stride + 1: stride]
        y += np.einsum('oc,bctn->botn', w[:, :, kk, 0], xs, optimize=True)
    return y


# --------------------------------------------------------------------------------------
# model/local_attention.py
# --------------------------------------------------------------------------------------
def semch_graph_conv(x, W, e, mask, bias=None):
    """SemCHGraphConv.forward, model/local_attention.py:35-53.  x: (B,T,J,Cin) -> (B,T,J,Cout).
    e: (Cout,nnz) scattered into the mask in row-major nonzero order (:25,:41);
    a shared-e (1,nnz) array restates model/sem_graph_conv.py:36-55."""
    Cout = W.shape[2]
    J = mask.shape[0]
    h0 = x @ W[0]                                    # :37
    h1 = x @ W[1]                                    # :38
    adj = np.full((Cout, J, J), -9e15, dtype=np.float32)   # :40
    rows, cols = np.nonzero(mask)
    adj[:, rows, cols] = np.broadcast_to(e, (Cout, len(rows)))   # :41
    adj = _softmax(adj, axis=2)                      # :42
    E = np.eye(J, dtype=np.float32)                  # :44-45
    a_self = adj * E
    a_nbr = adj * (1 - E)
    # :47  out[b,t,i,c] = sum_j a_self[c,i,j] h0[b,t,j,c] + a_nbr[c,i,j] h1[b,t,j,c]
    out = np.einsum('cij,btjc->btic', a_self, h0, optimize=True) + \
        np.einsum('cij,btjc->btic', a_nbr, h1, optimize=True)
    if bias is not None:
        out = out + bias.reshape(1, 1, 1, -1)
    return out.astype(np.float32)


def local_graph(x, p, prefix, masks, training=False, stats=None):
    """LocalGraph.forward, model/local_attention.py:130-151 (dropout = identity).
    x: (B,T,J,C) -> (B,T,J,C)."""
    sym, con = masks
    xs = semch_graph_conv(x, p[prefix + 'gcn_sym.W'], p[prefix + 'gcn_sym.e'], sym)   # :132
    ys = semch_graph_conv(x, p[prefix + 'gcn_con.W'], p[prefix + 'gcn_con.e'], con)   # :133
    xs = xs.transpose(0, 3, 1, 2)                    # :136
    ys = ys.transpose(0, 3, 1, 2)
    xs, s1 = batchnorm(xs, p, prefix + 'bn_1.', training)
    ys, s2 = batchnorm(ys, p, prefix + 'bn_2.', training)
    out = np.concatenate([relu(xs), relu(ys)], axis=1)   # :139-142
    out, s3 = batchnorm(conv1x1(out, p[prefix + 'cat_conv.weight']), p, prefix + 'cat_bn.', training)  # :143
    out = relu(out)                                  # :145-148
    if stats is not None:
        stats[prefix + 'bn_1.'] = s1
        stats[prefix + 'bn_2.'] = s2
        stats[prefix + 'cat_bn.'] = s3
    return out.transpose(0, 2, 3, 1)                 # :149


# --------------------------------------------------------------------------------------
# model/global_attention.py
# --------------------------------------------------------------------------------------
def global_graph(x, p, prefix):
    """GlobalGraph.forward, model/global_attention.py:52-82.  x: (BT, C, N) -> (BT, Cg, N).
    Literal: materialises concat_feature (BT, 2Ci, N, N) like :67-72."""
    wg = p[prefix + 'g.weight'][:, :, 0]
    wt = p[prefix + 'theta.weight'][:, :, 0]
    wp = p[prefix + 'phi.weight'][:, :, 0]
    g_x = np.einsum('oc,bcn->bon', wg, x, optimize=True) + p[prefix + 'g.bias'][None, :, None]     # :56
    g_x = g_x.transpose(0, 2, 1)                                                 # :57
    theta = np.einsum('oc,bcn->bon', wt, x, optimize=True) + p[prefix + 'theta.bias'][None, :, None]  # :60
    phi = np.einsum('oc,bcn->bon', wp, x, optimize=True) + p[prefix + 'phi.bias'][None, :, None]      # :62
    N = x.shape[2]
    theta_x = np.broadcast_to(theta[:, :, :, None], theta.shape + (N,))          # :67
    phi_x = np.broadcast_to(phi[:, :, None, :], phi.shape[:2] + (N, N))          # :68
    concat = np.concatenate([theta_x, phi_x], axis=1)                            # :71
    wc = p[prefix + 'concat_project.0.weight'].reshape(-1)
    f = np.einsum('c,bcij->bij', wc, concat, optimize=True)                      # :72
    att = np.where(f >= 0, f, np.float32(0.2) * f)                               # :74 LeakyReLU(0.2)
    att = _softmax(att.astype(np.float32), axis=-1) + p[prefix + 'C_k'][None]    # :76
    y = att @ g_x                                                                # :78
    return y.transpose(0, 2, 1).astype(np.float32)                               # :79-80


def multi_global_graph(x, p, prefix, training=False, stats=None):
    """MultiGlobalGraph.forward, model/global_attention.py:103-130 (dropout = identity).
    x: (B,T,J,C) -> (B,T,J,C)."""
    B, T, J, C = x.shape
    xf = x.reshape(B * T, J, C).transpose(0, 2, 1)                               # :105-109
    heads = 0
    while (prefix + 'attentions.%d.C_k' % heads) in p:
        heads += 1
    y = np.concatenate([global_graph(xf, p, prefix + 'attentions.%d.' % h) for h in range(heads)], axis=1)  # :111
    y = y.transpose(0, 2, 1).reshape(B, T, J, C)                                 # :114-118
    y = y.transpose(0, 3, 1, 2)                                                  # :121
    y, s = batchnorm(conv1x1(y, p[prefix + 'cat_conv.weight']), p, prefix + 'cat_bn.', training)   # :122
    y = relu(y)
    if stats is not None:
        stats[prefix + 'cat_bn.'] = s
    return y.transpose(0, 2, 3, 1)                                               # :128


# --------------------------------------------------------------------------------------
# model/gast_net.py
# --------------------------------------------------------------------------------------
def graph_attention_block(x, p, prefix, masks, training=False, stats=None):
    """GraphAttentionBlock.forward, model/gast_net.py:22-33.  x: (B,C,T,N) -> (B,2C,T,N)."""
    xl = x.transpose(0, 2, 3, 1)                                                 # :24
    a = local_graph(xl, p, prefix + 'local_graph_layer.', masks, training, stats)      # :26
    g = multi_global_graph(xl, p, prefix + 'global_graph_layer.', training, stats)    # :27
    cat = np.concatenate([xl, a, g], axis=-1).transpose(0, 3, 1, 2)              # :28-31
    y, s = batchnorm(conv1x1(cat, p[prefix + 'cat_conv.weight']), p, prefix + 'cat_bn.', training)  # :32
    if stats is not None:
        stats[prefix + 'cat_bn.'] = s
    return relu(y)


def model_geometry(filter_widths, causal, strided, dense=False):
    """pad / causal_shift / per-stage conv geometry, model/gast_net.py:136-155,213-232."""
    pad = [filter_widths[0] // 2]
    shift = [(filter_widths[0] // 2) if causal else 0]
    stages = []
    nd = filter_widths[0]
    for i in range(1, len(filter_widths)):
        pad.append((filter_widths[i] - 1) * nd // 2)
        if strided:
            shift.append((filter_widths[i] // 2) if causal else 0)
            stages.append(dict(k=filter_widths[i], stride=filter_widths[i], dilation=1))
        else:
            shift.append((filter_widths[i] // 2 * nd) if causal else 0)
            if dense:
                stages.append(dict(k=2 * pad[-1] + 1, stride=1, dilation=1))
            else:
                stages.append(dict(k=filter_widths[i], stride=1, dilation=nd))
        nd *= filter_widths[i]
    return pad, shift, stages


def forward(x, p, adj, filter_widths, causal=False, strided=False, dense=False,
            training=False, stats=None):
    """SpatioTemporalModel (strided=False, gast_net.py:159-177) or
    SpatioTemporalModelOptimized1f (strided=True, gast_net.py:236-251) forward incl.
    SpatioTemporalModelBase.forward (:84-104).  x: (B,T,J,F) -> (B,T_out,J,3).
    Dropout is the identity (eval, or train with p=0)."""
    masks = local_masks(adj)
    pad, shift, stages = model_geometry(filter_widths, causal, strided, dense)
    st = stats
    h = x.transpose(0, 3, 1, 2)                                                  # :162
    h, s = batchnorm(h, p, 'init_bn.', training)                                 # :163
    if st is not None:
        st['init_bn.'] = s
    h = conv_t(h, p['expand_conv.weight'], stride=(filter_widths[0] if strided else 1))
    h, s = batchnorm(h, p, 'expand_bn.', training)
    if st is not None:
        st['expand_bn.'] = s
    h = relu(h)                                                                  # :164
    h = graph_attention_block(h, p, 'layers_graph_conv.0.', masks, training, st)  # :165
    for i, sg in enumerate(stages):
        if strided:
            res = h[:, :, shift[i + 1] + filter_widths[i + 1] // 2:: filter_widths[i + 1]]   # :243
        else:
            res = h[:, :, pad[i + 1] + shift[i + 1]: h.shape[2] - pad[i + 1] + shift[i + 1]]  # :170
        h = conv_t(h, p['layers_conv.%d.weight' % (2 * i)], stride=sg['stride'], dilation=sg['dilation'])
        h, s = batchnorm(h, p, 'layers_bn.%d.' % (2 * i), training)
        if st is not None:
            st['layers_bn.%d.' % (2 * i)] = s
        h = relu(h)                                                              # :173
        h2 = conv1x1(h, p['layers_conv.%d.weight' % (2 * i + 1)])
        h2, s = batchnorm(h2, p, 'layers_bn.%d.' % (2 * i + 1), training)
        if st is not None:
            st['layers_bn.%d.' % (2 * i + 1)] = s
        h = res[:, :, :h2.shape[2]] + relu(h2)                                   # :174
        h = graph_attention_block(h, p, 'layers_graph_conv.%d.' % (i + 1), masks, training, st)  # :176
    y = conv1x1(h, p['shrink.weight'])                                           # :99
    return y.transpose(0, 2, 3, 1)                                               # :102


def mpjpe(pred, target):
    """common/loss.py:5-11."""
    assert pred.shape == target.shape
    return float(np.mean(np.linalg.norm(pred.astype(np.float64) - target.astype(np.float64), axis=-1)))


# --------------------------------------------------------------------------------------
# test-time augmentation around the forward
# --------------------------------------------------------------------------------------
def tta_prepare(seq, pad, causal_shift, kps_left, kps_right):
    """UnchunkedGenerator.next_epoch with augment=True, common/generators.py:210-233."""
    b = np.expand_dims(np.pad(seq, ((pad + causal_shift, pad - causal_shift), (0, 0), (0, 0)), 'edge'), axis=0)
    b = np.concatenate((b, b), axis=0)
    b[1, :, :, 0] *= -1
    b[1, :, kps_left + kps_right] = b[1, :, kps_right + kps_left]
    return b.astype(np.float32)


def tta_merge(pred, joints_left, joints_right):
    """main.py:314-318 / reconstruction.py:163-167."""
    p = pred.copy()
    p[1, :, :, 0] *= -1
    p[1, :, joints_left + joints_right] = p[1, :, joints_right + joints_left]
    return np.mean(p, axis=0, keepdims=True).squeeze(0)
