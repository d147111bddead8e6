"""Host glue between the drop-in nn.Module shells (model/*.py) and libgast_b200.so.

For each module instance (and device) the engine keeps one C handle.  The handle borrows
the module's own parameter/buffer storage (`gast_bind`, no copies), so optimiser steps and
`load_state_dict` are seen; derived constants are refreshed (`gast_prepare`) whenever a
tensor's (data_ptr, _version) signature changes.  torch provides device memory (outputs and
a cached workspace from the caching allocator) and the current CUDA stream -- plumbing only.

No fallback: CPU tensors, a missing library or an unsupported mode raise.
"""
import ctypes as C
import torch

from . import _lib as L

_FORCE_CORE = {'core': 0}   # 0 auto, 1 FFMA (A/B checks of the tcgen05 core on the GPU)


def set_gemm_core(core):
    """0 = auto (tcgen05 where the shape allows), 1 = FP32 FFMA core everywhere."""
    _FORCE_CORE['core'] = int(core)


class GastError(RuntimeError):
    pass


def _check(rc, what):
    if rc != 0:
        raise GastError('%s: %s' % (what, L.last_error()))


def _mask_lists(m):
    idx = m.nonzero()        # row-major order == order of `e` (local_attention.py:25,41)
    rows = (C.c_int32 * len(idx))(*[int(i) for i in idx[:, 0]])
    cols = (C.c_int32 * len(idx))(*[int(i) for i in idx[:, 1]])
    return rows, cols, len(idx)


def _m2d(m):
    m = m.detach().cpu()
    return m[0] if m.dim() == 3 else m


def named_tensors(module):
    """(key, tensor) pairs under the names `module.state_dict()` uses, in its order.  Unlike state_dict() this
    also works on the per-forward replicas of nn.DataParallel (trainval.py:56-61), whose weights are broadcast
    views kept as plain attributes (`_former_parameters`) instead of in `_parameters`."""
    out = []
    for prefix, mod in module.named_modules():
        pre = prefix + '.' if prefix else ''
        former = getattr(mod, '_former_parameters', None) or {}
        for k, v in mod._parameters.items():
            if v is not None:
                out.append((pre + k, v))
        for k, v in former.items():
            if v is not None and mod._parameters.get(k) is None:
                out.append((pre + k, v))
        for k, v in mod._buffers.items():
            if v is not None and k not in mod._non_persistent_buffers_set:
                out.append((pre + k, v))
    return out


def named_params(module):
    """named_parameters() that also sees the broadcast parameter views of a DataParallel replica
    (they are non-leaf tensors connected to the real parameters, so autograd routes the gradients)."""
    out = []
    for prefix, mod in module.named_modules():
        pre = prefix + '.' if prefix else ''
        former = getattr(mod, '_former_parameters', None) or {}
        for k, v in mod._parameters.items():
            if v is not None:
                out.append((pre + k, v))
        for k, v in former.items():
            if v is not None and mod._parameters.get(k) is None:
                out.append((pre + k, v))
    return out


class _HandleStore(dict):
    """device -> _Handle, kept in the module's __dict__ (so that the shallow-copied replicas of
    nn.DataParallel share it).  Native handles are per process and per device: a deepcopy / pickle of the
    module (EMA copies, torch.save(model)) gets an empty store and builds its own handles on first use."""

    def __deepcopy__(self, memo):
        return _HandleStore()

    def __reduce__(self):
        return (_HandleStore, ())


class _Handle(object):
    def __init__(self, module, kind, device, cfg_kw, sym=None, con=None):
        lib = L.load()
        self.lib = lib
        self.kind = kind
        self.device = device
        cfg = L.GastCfg()
        cfg.kind = kind
        cfg.device = device.index if device.index is not None else torch.cuda.current_device()
        for k, v in cfg_kw.items():
            if k == 'filter_widths':
                for i, fw in enumerate(v):
                    cfg.filter_widths[i] = int(fw)
                cfg.num_stages = len(v)
            else:
                setattr(cfg, k, int(v))
        keep = []
        if sym is not None:
            r, c, n = _mask_lists(_m2d(sym))
            cfg.sym_rows, cfg.sym_cols, cfg.sym_nnz = r, c, n
            keep += [r, c]
        if con is not None:
            r, c, n = _mask_lists(_m2d(con))
            cfg.con_rows, cfg.con_cols, cfg.con_nnz = r, c, n
            keep += [r, c]
        self.h = C.c_void_p()
        _check(lib.gast_create(C.byref(self.h), C.byref(cfg)), 'gast_create')
        self.sig = None
        self.ws = None
        self.core = None

    def __del__(self):
        try:
            if getattr(self, 'h', None) and self.h.value:
                self.lib.gast_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass

    def refresh(self, module, stream):
        """(re)bind + prepare when any parameter/buffer moved or changed."""
        items = named_tensors(module)
        sig = tuple((v.data_ptr(), v._version) for _, v in items)
        if _FORCE_CORE['core'] != self.core:
            _check(self.lib.gast_set_gemm_core(self.h, _FORCE_CORE['core']), 'gast_set_gemm_core')
            self.core = _FORCE_CORE['core']
        if sig == self.sig:
            return
        n = len(items)
        keys = (C.c_char_p * n)()
        ptrs = (C.c_void_p * n)()
        numel = (C.c_int64 * n)()
        for i, (k, v) in enumerate(items):
            if not v.is_cuda or v.device != self.device:
                raise GastError('parameter %r is on %s, expected %s (call .cuda() on the module)'
                                % (k, v.device, self.device))
            if v.dtype == torch.float32 and not v.is_contiguous():
                raise GastError('parameter %r is not contiguous' % k)
            keys[i] = k.encode()
            ptrs[i] = v.data_ptr()
            numel[i] = v.numel()
        _check(self.lib.gast_bind(self.h, n, keys, ptrs, numel), 'gast_bind')
        _check(self.lib.gast_prepare(self.h, C.c_void_p(stream)), 'gast_prepare')
        self.sig = sig

    def bind_only(self, module):
        """bind parameter/buffer storage without the eval-mode prepare (training path)"""
        items = named_tensors(module)
        n = len(items)
        keys = (C.c_char_p * n)()
        ptrs = (C.c_void_p * n)()
        numel = (C.c_int64 * n)()
        for i, (k, v) in enumerate(items):
            if not v.is_cuda or v.device != self.device:
                raise GastError('parameter %r is on %s, expected %s' % (k, v.device, self.device))
            if not v.is_contiguous():
                raise GastError('parameter %r is not contiguous' % k)
            keys[i] = k.encode()
            ptrs[i] = v.data_ptr()
            numel[i] = v.numel()
        _check(self.lib.gast_bind(self.h, n, keys, ptrs, numel), 'gast_bind')
        self.sig = None

    def forward(self, x, y, B, T, strided_now, stream):
        need = self.lib.gast_workspace_bytes(self.h, B, T, strided_now)
        if need == 0:
            raise GastError('gast_workspace_bytes: %s' % L.last_error())
        if self.ws is None or self.ws.numel() < need:
            self.ws = None
            self.ws = torch.empty(int(need), dtype=torch.uint8, device=self.device)
        _check(self.lib.gast_forward(self.h, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), B, T,
                                     strided_now, C.c_void_p(self.ws.data_ptr()), self.ws.numel(),
                                     C.c_void_p(stream)), 'gast_forward')

    def forward_mpjpe(self, x, target, y, loss, B, T, strided_now, stream):
        need = self.lib.gast_workspace_bytes(self.h, B, T, strided_now)
        if need == 0:
            raise GastError('gast_workspace_bytes: %s' % L.last_error())
        if self.ws is None or self.ws.numel() < need:
            self.ws = None
            self.ws = torch.empty(int(need), dtype=torch.uint8, device=self.device)
        _check(self.lib.gast_forward_mpjpe(self.h, C.c_void_p(x.data_ptr()), C.c_void_p(target.data_ptr()),
                                           C.c_void_p(y.data_ptr()), C.c_void_p(loss.data_ptr()), B, T, strided_now,
                                           C.c_void_p(self.ws.data_ptr()), self.ws.numel(), C.c_void_p(stream)),
               'gast_forward_mpjpe')

    def launches(self):
        return int(self.lib.gast_last_launch_count(self.h))


def _require_cuda(x, what):
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise GastError('%s: input must be a CUDA tensor -- this implementation has no CPU path' % what)
    if x.dtype != torch.float32:
        raise GastError('%s: input must be float32 (got %s)' % (what, x.dtype))


def _handle_for(module, device, factory):
    store = module.__dict__.get('_gast_handles')
    if store is None:
        store = module.__dict__.setdefault('_gast_handles', _HandleStore())
    key = (device.type, device.index)
    h = store.get(key)
    if h is None:
        h = factory()
        store[key] = h
    return h


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


class _TrainFn(torch.autograd.Function):
    """Training-mode forward/backward of the whole model on the CUDA library (gast_forward_train /
    gast_backward).  Parameters are passed as inputs so that autograd routes their gradients."""

    @staticmethod
    def forward(ctx, module, handle, x, names, *params):
        dev = x.device
        B, T = int(x.shape[0]), int(x.shape[1])
        lib = handle.lib
        with torch.cuda.device(dev):
            st = _stream(dev)
            handle.bind_only(module)
            need = lib.gast_train_workspace_bytes(handle.h, B, T, C.c_float(float(module._gast_dropout)))
            if need == 0:
                raise GastError('gast_train_workspace_bytes: %s' % L.last_error())
            ws = torch.empty(int(need), dtype=torch.uint8, device=dev)
            T_out = lib.gast_out_frames(handle.h, T, 1 if module._gast_strided else 0)
            if T_out <= 0:
                raise GastError('forward: %s' % L.last_error())
            y = torch.empty((B, T_out, module.num_joints_in, 3), dtype=torch.float32, device=dev)
            p = float(module._gast_dropout)
            seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p > 0 else 0
            # device-side dropout counter (set by a graph-capturing trainer): keeps the masks fresh across replays
            ds = module.__dict__.get('_gast_dropout_state')
            _check(lib.gast_set_dropout_state(handle.h, C.c_void_p(ds.data_ptr() if ds is not None else None)),
                   'gast_set_dropout_state')
            _check(lib.gast_forward_train(handle.h, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), B, T,
                                          C.c_float(p), C.c_uint64(seed), C.c_void_p(ws.data_ptr()), ws.numel(),
                                          C.c_void_p(st)), 'gast_forward_train')
        # BatchNorm bookkeeping that torch does in Python: num_batches_tracked (this also bumps the
        # buffers' versions, so the eval-mode constants are refreshed on the next eval forward)
        with torch.no_grad():
            nbt = [mod.num_batches_tracked for mod in module.modules() if isinstance(mod, torch.nn.BatchNorm2d)]
            if nbt:
                torch._foreach_add_(nbt, 1)              # one launch instead of one per BatchNorm
        handle.sig = None
        ctx.handle, ctx.ws, ctx.names, ctx.params = handle, ws, names, params
        ctx.direct = bool(module.__dict__.get('_gast_direct_grads', False))
        ctx.x = x            # the backward re-reads the input (init_bn statistics): keep it alive
        ctx.dev = dev
        return y

    @staticmethod
    def backward(ctx, dy):
        handle, ws, names, params = ctx.handle, ctx.ws, ctx.names, ctx.params
        lib = handle.lib
        total = sum(p.numel() for p in params)
        n = len(params)
        keys = (C.c_char_p * n)()
        ptrs = (C.c_void_p * n)()
        numel = (C.c_int64 * n)()
        # A trainer that owns the gradient buffers (module._gast_direct_grads, set by gast_b200.trainer: every .grad is
        # a view of its flat all-reduce buffer and one backward runs per step) gets the gradients WRITTEN straight into
        # them: autograd's AccumulateGrad would otherwise launch one add per parameter (166 per step at 27f/128ch,
        # 0.46 ms of a 8.3 ms step).  Everyone else gets them through autograd as usual.
        direct = ctx.direct and all(p.is_leaf and p.grad is not None and p.grad.is_contiguous() and
                                    p.grad.dtype == torch.float32 and p.grad.device == p.device and
                                    p.grad.shape == p.shape for p in params)
        flat = None if direct else torch.empty(total, dtype=torch.float32, device=ctx.dev)
        views, off = [], 0
        for i, (k, p) in enumerate(zip(names, params)):
            v = p.grad if direct else flat[off:off + p.numel()].view_as(p)
            views.append(v)
            keys[i] = k.encode()
            ptrs[i] = v.data_ptr()
            numel[i] = p.numel()
            off += p.numel()
        dy = dy.contiguous()
        with torch.cuda.device(ctx.dev):
            _check(lib.gast_bind_grads(handle.h, n, keys, ptrs, numel), 'gast_bind_grads')
            _check(lib.gast_backward(handle.h, C.c_void_p(dy.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(),
                                     C.c_void_p(_stream(ctx.dev))), 'gast_backward')
        ctx.ws = None
        ctx.x = None
        if direct:
            return (None, None, None, None) + (None,) * n
        return (None, None, None, None) + tuple(views)


def _no_train(module, what):
    if module.training:
        raise GastError('%s: the training-mode forward/backward exists for the whole model only '
                        '(SpatioTemporalModel / SpatioTemporalModelOptimized1f, what main.train() drives); a sub-module '
                        'on its own runs in eval mode -- call .eval() (there is no silent fallback)' % what)


# ------------------------------------------------------------------------------------------
def model_handle(module, dev):
    """the C handle of a SpatioTemporalModel / ...Optimized1f instance on `dev` (created on first use)"""
    blk0 = module.layers_graph_conv[0].local_graph_layer

    def make():
        return _Handle(module, L.KIND_MODEL, dev, dict(
            num_joints=module.num_joints_in, in_features=module.in_features,
            channels=module._gast_channels, filter_widths=list(module.filter_widths),
            causal=module._gast_causal, dense=module._gast_dense, strided=module._gast_strided, heads=4),
            sym=blk0.gcn_sym.m, con=blk0.gcn_con.m)
    return _handle_for(module, dev, make)


def run_model(module, x):
    """SpatioTemporalModelBase.forward (gast_net.py:84-104) on the CUDA library."""
    _require_cuda(x, 'SpatioTemporalModel.forward')
    dev = x.device
    h = model_handle(module, dev)
    x = x.contiguous()
    B, T = int(x.shape[0]), int(x.shape[1])
    if module.training:
        # both of the reference's training models: Optimized1f (main.py:166-170) and, with
        # --disable-optimizations / stride > 1, the dilated SpatioTemporalModel (main.py:171-175)
        np_ = named_params(module)
        names = [k for k, _ in np_]
        params = [p for _, p in np_]
        y = _TrainFn.apply(module, h, x, names, *params)
        module.__dict__['_gast_last_launches'] = h.launches()
        return y
    # A dilated model fed exactly one receptive field computes only what the strided
    # schedule computes (same arithmetic per output); skip the unused positions.
    strided_now = 1 if (module._gast_strided or
                        (T == module.receptive_field() and not module._gast_dense)) else 0
    with torch.cuda.device(dev):
        st = _stream(dev)
        h.refresh(module, st)
        T_out = h.lib.gast_out_frames(h.h, T, strided_now)
        if T_out <= 0:
            raise GastError('forward: %s' % L.last_error())
        y = torch.empty((B, T_out, module.num_joints_in, 3), dtype=torch.float32, device=dev)
        h.forward(x, y, B, T, strided_now, st)
    module.__dict__['_gast_last_launches'] = h.launches()
    return y


def run_model_mpjpe(module, x, target):
    """`predicted = model(x); error = mpjpe(predicted, target)` (main.py:270-300, common/loss.py:5-11) as ONE library
    call: the loss is taken in the shrink kernel's epilogue.  Eval mode; returns (predicted, loss) with loss a 0-d
    CUDA tensor.  (The training loss goes through gast_b200.pipeline.mpjpe, which also produces the gradient.)"""
    _require_cuda(x, 'run_model_mpjpe')
    _require_cuda(target, 'run_model_mpjpe (target)')
    _no_train(module, 'run_model_mpjpe')
    dev = x.device
    h = model_handle(module, dev)
    x = x.contiguous()
    B, T = int(x.shape[0]), int(x.shape[1])
    strided_now = 1 if (module._gast_strided or
                        (T == module.receptive_field() and not module._gast_dense)) else 0
    with torch.cuda.device(dev):
        st = _stream(dev)
        h.refresh(module, st)
        T_out = h.lib.gast_out_frames(h.h, T, strided_now)
        if T_out <= 0:
            raise GastError('forward: %s' % L.last_error())
        shape = (B, T_out, module.num_joints_in, 3)
        if tuple(target.shape) != shape or target.device != dev:
            raise GastError('run_model_mpjpe: target must be %s on %s (got %s on %s)'
                            % (shape, dev, tuple(target.shape), target.device))
        target = target.contiguous()
        y = torch.empty(shape, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        h.forward_mpjpe(x, target, y, loss, B, T, strided_now, st)
    module.__dict__['_gast_last_launches'] = h.launches()
    return y, loss


def run_block(module, x):
    """GraphAttentionBlock.forward: x (B,C,T,N) -> (B,2C,T,N) (gast_net.py:22-33)."""
    _require_cuda(x, 'GraphAttentionBlock.forward')
    _no_train(module, 'GraphAttentionBlock.forward')
    dev = x.device
    lg = module.local_graph_layer
    Bn, Cc, T, N = x.shape

    def make():
        return _Handle(module, L.KIND_BLOCK, dev, dict(num_joints=N, channels=Cc, heads=4),
                       sym=lg.gcn_sym.m, con=lg.gcn_con.m)
    h = _handle_for(module, dev, make)
    xl = x.permute(0, 2, 3, 1).contiguous()
    with torch.cuda.device(dev):
        st = _stream(dev)
        h.refresh(module, st)
        y = torch.empty((Bn, T, N, 2 * Cc), dtype=torch.float32, device=dev)
        h.forward(xl, y, Bn * T, 1, 0, st)
    return y.permute(0, 3, 1, 2)


def run_local(module, x):
    """LocalGraph.forward: (B,T,J,C) -> (B,T,J,C) (local_attention.py:130-151)."""
    _require_cuda(x, 'LocalGraph.forward')
    _no_train(module, 'LocalGraph.forward')
    dev = x.device
    Bn, T, N, Cc = x.shape

    def make():
        return _Handle(module, L.KIND_LOCAL, dev, dict(num_joints=N, channels=Cc),
                       sym=module.gcn_sym.m, con=module.gcn_con.m)
    h = _handle_for(module, dev, make)
    xl = x.contiguous()
    with torch.cuda.device(dev):
        st = _stream(dev)
        h.refresh(module, st)
        y = torch.empty((Bn, T, N, Cc), dtype=torch.float32, device=dev)
        h.forward(xl, y, Bn * T, 1, 0, st)
    return y


def run_semch(module, x):
    """SemCHGraphConv.forward / SemGraphConv.forward: (B,T,J,Cin) -> (B,T,J,Cout)."""
    _require_cuda(x, 'SemCHGraphConv.forward')
    dev = x.device
    Bn, T, N, Cc = x.shape
    shared = 1 if module.e.shape[0] == 1 and module.out_features != 1 else 0

    def make():
        return _Handle(module, L.KIND_SEMCH, dev, dict(
            num_joints=N, channels=Cc, channels_out=module.out_features, semch_shared_e=shared,
            semch_bias=1 if module.bias is not None else 0), sym=module.m)
    h = _handle_for(module, dev, make)
    xl = x.contiguous()
    with torch.cuda.device(dev):
        st = _stream(dev)
        h.refresh(module, st)
        y = torch.empty((Bn, T, N, module.out_features), dtype=torch.float32, device=dev)
        h.forward(xl, y, Bn * T, 1, 0, st)
    return y


def run_multi_global(module, x):
    """MultiGlobalGraph.forward: (B,T,J,C) -> (B,T,J,C) (global_attention.py:103-130)."""
    _require_cuda(x, 'MultiGlobalGraph.forward')
    _no_train(module, 'MultiGlobalGraph.forward')
    dev = x.device
    Bn, T, N, Cc = x.shape

    def make():
        return _Handle(module, L.KIND_MGLOBAL, dev, dict(num_joints=N, channels=Cc,
                                                         heads=module.num_non_local))
    h = _handle_for(module, dev, make)
    xl = x.contiguous()
    with torch.cuda.device(dev):
        st = _stream(dev)
        h.refresh(module, st)
        y = torch.empty((Bn, T, N, Cc), dtype=torch.float32, device=dev)
        h.forward(xl, y, Bn * T, 1, 0, st)
    return y


def run_global_head(module, x):
    """GlobalGraph.forward: (B*T, C, N) -> (B*T, Cg, N) (global_attention.py:52-82)."""
    _require_cuda(x, 'GlobalGraph.forward')
    dev = x.device
    BT, Cc, N = x.shape
    if module.g_channels != module.inter_channels:
        raise GastError('GlobalGraph with inter_channels == in_channels//2 is not on the lifting path')

    def make():
        return _Handle(module, L.KIND_GLOBAL_HEAD, dev, dict(num_joints=N, channels=Cc,
                                                             channels_out=module.inter_channels, heads=1))
    h = _handle_for(module, dev, make)
    xl = x.permute(0, 2, 1).contiguous()
    with torch.cuda.device(dev):
        st = _stream(dev)
        h.refresh(module, st)
        y = torch.empty((BT, N, module.g_channels), dtype=torch.float32, device=dev)
        h.forward(xl, y, BT, 1, 0, st)
    return y.permute(0, 2, 1)


LAUNCH_KINDS = ['expand', 'gemm_ffma_plain', 'gemm_ffma_semch', 'gemm_ffma_global', 'rowdot', 'shrink',
                'gemm_tc_plain', 'gemm_tc_semch', 'gemm_tc_global', 'global_mix']


def profile_forward(module, x, reps=5):
    """Per-launch device times of the model forward (CUDA events inside the library, on the
    launching stream).  Returns per-kernel-kind ms per step and the GEMM-family totals."""
    dev = x.device
    with torch.no_grad():
        run_model(module, x)
    h = module.__dict__['_gast_handles'][(dev.type, dev.index)]
    lib = h.lib
    _check(lib.gast_set_timing(h.h, 1), 'gast_set_timing')
    acc = {}
    n_launch = {}
    try:
        for _ in range(reps):
            with torch.no_grad():
                run_model(module, x)
            ms = (C.c_float * 512)()
            kinds = (C.c_int32 * 512)()
            n = lib.gast_get_timings(h.h, 512, ms, kinds)
            if n < 0:
                raise GastError('gast_get_timings failed')
            n_launch = {}
            for i in range(n):
                k = LAUNCH_KINDS[kinds[i]]
                acc[k] = acc.get(k, 0.0) + ms[i]
                n_launch[k] = n_launch.get(k, 0) + 1
    finally:
        lib.gast_set_timing(h.h, 0)
    per = {k: v / reps for k, v in acc.items()}
    # the GEMM family = every fused channel-contraction launch, plus the attention mix where it runs as its own
    # kernel (it is the epilogue of the `g` GEMM in the fused form)
    fam = lambda k: k.startswith('gemm') or k == 'global_mix'
    gemm = sum(v for k, v in per.items() if fam(k))
    ng = sum(v for k, v in n_launch.items() if fam(k))
    tc = int(lib.gast_last_tc_launch_count(h.h))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        e0.record()
        run_model(module, x)
        e1.record()
    torch.cuda.synchronize(dev)
    return {'per_kernel_ms': per, 'launches': n_launch, 'gemm_ms_per_step': gemm, 'gemm_launches': ng,
            'step_ms': e0.elapsed_time(e1), 'workspace_bytes': int(h.ws.numel()) if h.ws is not None else 0,
            'gemm_core': ('tcgen05-3xtf32 x%d + ffma x%d + attention-mix x%d'
                          % (tc, ng - tc - n_launch.get('global_mix', 0), n_launch.get('global_mix', 0))) if tc else 'ffma-fp32'}
