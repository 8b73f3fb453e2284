"""Forward/backward engine of the DeepCAM network on top of a layer-level backend.

The nn.Modules in architecture/deeplab_xception.py only own parameters (so state_dict, optimizers and
checkpoints behave exactly as in the reference).  Their forward() describes the network to this engine as a
sequence of layer operations on channels-last activations; the engine executes them immediately on the
backend and, when gradients are needed, records one closure per operation.  backward() replays the closures
in reverse with explicit gradient-buffer management: the first writer of a gradient buffer writes, later
writers accumulate in their epilogue (no separate add kernels), concat buffers are never materialised twice
(producers write channel slices), and parameter gradients land in one flat fp32 buffer that the
data-parallel wrapper all-reduces bucket by bucket while backward is still running.

The whole plan is wrapped in ONE torch.autograd.Function per root module call, so autograd sees
(input, *parameters) -> outputs and the reference training loop (TR:345-371) runs unchanged.
"""
import os
import threading

import torch

from .backend import BnSpec, ConvSpec, CudaBackend, DwSpec  # noqa: F401  (re-exported for the modules)

_PRECISIONS = {"bf16": torch.bfloat16, "fp32": torch.float32}
_state = threading.local()


def default_precision():
    return os.environ.get("DEEPCAM_B200_PRECISION", "bf16")


_deterministic_override = None


def set_deterministic(flag):
    """True / False: force the deterministic kernels on / off; None: follow DEEPCAM_B200_DETERMINISTIC=1 and
    torch.use_deterministic_algorithms() (what a user of the reference would call to get reproducible cuDNN gradients)."""
    global _deterministic_override
    _deterministic_override = None if flag is None else bool(flag)


def deterministic():
    """Whether module calls use the two-stage (workspace + ordered reduction) weight-gradient kernels and the single-writer
    pooling reduction: bit-identical gradients from run to run, at the cost of one extra launch per split weight gradient."""
    if _deterministic_override is not None:
        return _deterministic_override
    if os.environ.get("DEEPCAM_B200_DETERMINISTIC", "0") not in ("0", "false", "False", ""):
        return True
    return bool(torch.are_deterministic_algorithms_enabled())


def set_backend_factory(factory):
    """TEST HOOK: tests/ installs a torch-CPU interpreter of the backend interface to check the graph logic
    without a GPU.  The product never calls this; the default factory builds the CUDA backend and raises on CPU."""
    _state.factory = factory


def _make_backend(precision, device):
    factory = getattr(_state, "factory", None)
    if factory is not None:
        return factory(_PRECISIONS[precision], device)
    if device.type != "cuda":
        raise RuntimeError(
            "deepcam_b200 runs on CUDA (sm_100a) only: input is on %s and there is no CPU fallback. "
            "Move the module and its inputs to a B200 device." % device)
    return CudaBackend(_PRECISIONS[precision], device)


class Act:
    """Channels-last activation: tensor [N,H,W,C] (possibly a channel slice of a concat buffer) + its gradient."""
    __slots__ = ("t", "_grad", "is_relu", "needs_grad", "parent", "c_off", "bn_sums", "consumed", "bn_src", "bn_reduced", "deferred")

    def __init__(self, t, is_relu=False, needs_grad=True, parent=None, c_off=0):
        self.t = t
        self.deferred = None      # (x, ConvSpec, shape, dtype): a convolution not launched yet (eval-mode BatchNorm folding)
        self.bn_sums = None       # (BatchNorm module, workspace) when the producing conv already accumulated the batch sums
        self.consumed = 0         # operations that have taken this activation as an input so far (forward order)
        self.bn_src = None        # (y, spec, sums, relu, residual) when this is the output of a train-mode BatchNorm
        self.bn_reduced = None    # backward workspace when the last gradient writer already masked + reduced (fused dw bwd)
        self._grad = None
        self.is_relu = is_relu
        self.needs_grad = needs_grad
        self.parent = parent
        self.c_off = c_off

    @property
    def shape(self):
        if self.t is None:
            return self.deferred[2]
        return tuple(self.t.shape)

    @property
    def grad(self):
        if self.parent is not None:
            g = self.parent.grad
            return None if g is None else g[..., self.c_off:self.c_off + self.t.shape[3]]
        return self._grad

    @grad.setter
    def grad(self, g):
        if self.parent is not None:
            raise RuntimeError("gradient of a concat slice is owned by the concat buffer")
        self._grad = g

    def slice(self, c_off, c):
        return Act(self.t[..., c_off:c_off + c], parent=self, c_off=c_off)


class GradStore:
    """Flat fp32 gradient buffer with one parameter-shaped view per parameter (gradient-as-bucket-view).

    Views handed to autograd are fresh tensors, so AccumulateGrad adopts them without a copy when the
    parameter's .grad is None (optimizer.zero_grad() default).  If any parameter still holds a gradient that
    aliases the buffer (gradient accumulation, zero_grad(set_to_none=False)) a new buffer is allocated for the
    next backward so the held gradients stay intact and autograd accumulates into them."""

    def __init__(self, params, device):
        self.params = list(params)
        self.offsets = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4          # keep every view 16-byte aligned
        self.total = max(off, 4)
        self.device = device
        self.flat = None
        self._index = {id(p): i for i, p in enumerate(self.params)}
        self._views = {}
        self.on_ready = None      # callback(param_index): all gradient kernels of that parameter have been launched

    def matches(self, params, device):
        return self.device == device and len(self.params) == len(params) and all(a is b for a, b in zip(self.params, params))

    def begin_backward(self):
        """Pick (or allocate) the flat buffer for this backward; returns it un-zeroed."""
        if self.flat is not None:
            lo = self.flat.data_ptr()
            hi = lo + self.flat.numel() * 4
            for p in self.params:
                g = p.grad
                if g is not None and lo <= g.data_ptr() < hi:
                    self.flat = None
                    break
        if self.flat is None:
            self.flat = torch.empty(self.total, dtype=torch.float32, device=self.device)
        self._views = {}
        return self.flat

    def view(self, p):
        v = self._views.get(id(p))
        if v is None:
            o = self.offsets[self._index[id(p)]]
            v = self.flat[o:o + p.numel()].view(p.shape)
            self._views[id(p)] = v
        return v

    def take_views(self):
        """Parameter-ordered gradient views for autograd; drops our own references so autograd can adopt them."""
        out = tuple(self.view(p) if p.requires_grad else None for p in self.params)
        self._views = {}
        return out

    def ready(self, p):
        if self.on_ready is not None:
            self.on_ready(self._index[id(p)])


class Engine:
    def __init__(self, backend, record, grads=None):
        self.be = backend
        self.record = record
        self.grads = grads
        self.tape = []
        self.bn_trained = []        # BatchNorm modules that ran in training mode (num_batches_tracked += 1)

    # ---- activations ---------------------------------------------------------------------------------
    def new_act(self, n, h, w, c, dtype=None):
        return Act(self.be.empty(n, h, w, c, dtype))

    def from_nchw(self, x, needs_grad, c_pad=None):
        return Act(self.be.from_nchw(x, c_pad), needs_grad=needs_grad)

    def grad_target(self, act):
        """(gradient tensor, accumulate?) for a writer of act's gradient: the first writer writes."""
        g = act.grad
        if g is not None:
            return g, True
        if act.parent is not None:
            raise RuntimeError("concat-slice gradient requested before the concat consumer ran backward")
        n, h, w, c = act.shape
        act.grad = self.be.empty(n, h, w, c, act.t.dtype)
        return act.grad, False

    # ---- operations -------------------------------------------------------------------------------------
    def conv(self, x, spec, out=None, out_c=None, out_dtype=None, bn=None):
        """bn: the BnSpec that will normalise the result next (Engine.bn on the returned Act): in training mode the
        convolution kernel then also produces the batch sums, see CudaBackend.conv_fwd."""
        self._materialize(x)
        n, h, w, _ = x.shape
        x.consumed += 1
        ho, wo = spec.out_hw(h, w)
        want = bn is not None and (bool(bn.module.training) or bn.module.running_mean is None)
        if (bn is not None and not want and not self.record and out is None and out_c is None
                and hasattr(self.be, "conv_bn_eval_fwd")):
            # eval mode, nothing recorded: do not launch yet - Engine.bn folds the normalisation into this GEMM's epilogue
            # (anything else that touches the result first materialises it, see _materialize)
            act = Act(None)
            act.deferred = (x, spec, (n, ho, wo, spec.co), out_dtype or x.t.dtype)
            return act
        if out is None:
            out = self.new_act(n, ho, wo, out_c or spec.co, out_dtype or x.t.dtype)
        elif out.shape[:3] != (n, ho, wo):
            raise RuntimeError("conv %s: output buffer %s does not match %s" % (spec.name, out.shape, (n, ho, wo)))
        if want:
            out.bn_sums = (bn.module, self.be.conv_fwd(x.t, spec, out.t, want_bn_sums=True))
        else:
            self.be.conv_fwd(x.t, spec, out.t)
        if self.record:
            self.tape.append(lambda: self._conv_bwd(x, out, spec))
        return out

    def _materialize(self, act):
        """Launch a deferred convolution as it is (its consumer turned out not to be a foldable eval-mode BatchNorm)."""
        if act.deferred is not None:
            x, spec, shape, dtype = act.deferred
            act.deferred = None
            act.t = self.be.empty(*shape, dtype)
            self.be.conv_fwd(x.t, spec, act.t)
        return act

    def _conv_bwd(self, x, out, spec):
        dy = out.grad
        if dy is None:
            return
        if spec.weight.requires_grad:
            bgrad = self.grads.view(spec.bias) if (spec.bias is not None and spec.bias.requires_grad) else None
            with self.be.side_branch():
                self.be.conv_bwd_weight(x.t, dy, spec, self.grads.view(spec.weight), bgrad)
            self.grads.ready(spec.weight)
            if bgrad is not None:
                self.grads.ready(spec.bias)
        if x.needs_grad:
            dx, acc = self.grad_target(x)
            self.be.conv_bwd_data(dy, spec, dx, acc)

    def dw(self, x, spec):
        self._materialize(x)
        n, h, w, c = x.shape
        ho, wo = spec.out_hw(h, w)
        out = self.new_act(n, ho, wo, c, x.t.dtype)
        first = x.consumed == 0       # first consumer in forward order = LAST writer of x's gradient in backward
        x.consumed += 1
        self.be.dw_fwd(x.t, spec, out.t)
        if self.record:
            self.tape.append(lambda: self._dw_bwd(x, out, spec, first))
        return out

    def _dw_bwd(self, x, out, spec, last_writer=False):
        dy = out.grad
        if dy is None:
            return
        if spec.weight.requires_grad:
            with self.be.side_branch():
                self.be.dw_bwd_weight(x.t, dy, spec, self.grads.view(spec.weight))
            self.grads.ready(spec.weight)
        if x.needs_grad:
            dx, acc = self.grad_target(x)
            src = x.bn_src
            if last_writer and src is not None and x.parent is None and hasattr(self.be, "dw_bwd_data_bnred"):
                # x = relu(bn(y) [+ residual]) and this is the last contribution to its gradient: mask it and reduce it for
                # BatchNorm backward right here; _bn_bwd then only runs the element-wise pass
                y, _bspec, sums, relu, residual = src
                act = x.t if (relu and residual is not None) else None
                rws = self.be.dw_bwd_data_bnred(dy, spec, dx, acc, y.t, act, sums, relu)
                if rws is not None:
                    x.bn_reduced = rws
                    return
            self.be.dw_bwd_data(dy, spec, dx, acc)

    def bn(self, y, spec, relu, residual=None, out=None):
        """out = [relu](bn(y) [+ residual]); spec None = identity (plain relu / add / copy)."""
        if y.deferred is not None:
            x_in, cspec, shape, dtype = y.deferred
            if spec is not None and residual is None and (out is None or out.shape == shape):
                dst = out if out is not None else self.new_act(*shape, dtype)
                if self.be.conv_bn_eval_fwd(x_in.t, cspec, spec, relu, dst.t):
                    y.deferred = None
                    y.consumed += 1
                    dst.is_relu = relu
                    return dst
            self._materialize(y)
        if residual is not None:
            self._materialize(residual)
        n, h, w, c = y.shape
        y.consumed += 1
        if residual is not None:
            residual.consumed += 1
        if out is None:
            out = self.new_act(n, h, w, c, y.t.dtype)
        elif out.shape != y.shape:
            raise RuntimeError("output buffer %s does not match the normalised tensor %s (concat size mismatch)"
                               % (out.shape, y.shape))
        out.is_relu = relu
        training = bool(spec.module.training) if spec is not None else False
        if spec is not None and not training and spec.module.running_mean is None:
            training = True          # track_running_stats=False always uses batch statistics
        ready = None
        pre = getattr(y, "bn_sums", None)
        if pre is not None and spec is not None and training and pre[0] is spec.module and pre[1] is not None:
            ready = pre[1]
            y.bn_sums = None
        if ready is not None:
            sums = self.be.bn_fwd(y.t, spec, relu, residual.t if residual is not None else None, out.t, training, ready_sums=ready)
        else:
            sums = self.be.bn_fwd(y.t, spec, relu, residual.t if residual is not None else None, out.t, training)
        if spec is not None and training and spec.module.num_batches_tracked is not None:
            self.bn_trained.append(spec.module)
        if self.record:
            if spec is not None and training and sums is not None and out.parent is None and out.t.dtype == y.t.dtype:
                out.bn_src = (y, spec, sums, relu, residual)
            self.tape.append(lambda: self._bn_bwd(y, out, spec, sums, relu, residual, training))
        return out

    def bn_dw(self, y, spec, relu, dwspec):
        """(a, t) with a = [relu](bn(y)) and t = depthwise(a): the ReLU -> SeparableConv2d_same chain of a Block (DX:79-97).
        One launch when the backend can apply the BatchNorm while the depthwise kernel loads its tile (train mode, batch sums
        already produced by the GEMM epilogue, stride 1, dilation 1); otherwise exactly bn() followed by dw().  The tape gets
        the same two backward closures either way."""
        if y.deferred is not None:                            # eval mode: conv + BatchNorm(+ReLU) fold, then a plain depthwise
            a = self.bn(y, spec, relu)
            return a, self.dw(a, dwspec)
        training = bool(spec.module.training) or spec.module.running_mean is None
        pre = getattr(y, "bn_sums", None)
        ready = pre[1] if (pre is not None and training and pre[0] is spec.module and pre[1] is not None) else None
        if ready is not None and hasattr(self.be, "bn_dw_fwd"):
            n, h, w, c = y.shape
            a = self.new_act(n, h, w, c, y.t.dtype)
            t = self.new_act(n, h, w, c, y.t.dtype)
            sums = self.be.bn_dw_fwd(y.t, spec, relu, dwspec, a.t, t.t, ready)
            if sums is not None:
                y.bn_sums = None
                y.consumed += 1
                a.is_relu = relu
                a.consumed += 1
                if spec.module.num_batches_tracked is not None:
                    self.bn_trained.append(spec.module)
                if self.record:
                    a.bn_src = (y, spec, sums, relu, None)
                    self.tape.append(lambda: self._bn_bwd(y, a, spec, sums, relu, None, True))
                    self.tape.append(lambda: self._dw_bwd(a, t, dwspec, True))
                return a, t
        a = self.bn(y, spec, relu)
        return a, self.dw(a, dwspec)

    def _bn_bwd(self, y, out, spec, sums, relu, residual, training):
        dout = out.grad
        if dout is None:
            return
        dres, racc = (None, False)
        if residual is not None and residual.needs_grad:
            dres, racc = self.grad_target(residual)
        dy = None
        if y.needs_grad:
            dy, yacc = self.grad_target(y)
            if yacc:
                raise RuntimeError("BatchNorm input gradient must have a single writer")
        dgamma = dbeta = None
        if spec is not None:
            m = spec.module
            if m.weight is not None and m.weight.requires_grad:
                dgamma = self.grads.view(m.weight)
            if m.bias is not None and m.bias.requires_grad:
                dbeta = self.grads.view(m.bias)
        rws = out.bn_reduced
        if rws is not None:         # the gradient arrived masked, with its BatchNorm sums (fused depthwise backward)
            out.bn_reduced = None
            self.be.bn_bwd_reduced(dout, y.t, spec, sums, rws, dy, dres, racc, dgamma, dbeta)
        else:
            self.be.bn_bwd(dout, out.t, y.t, spec, sums, relu, dy, dres, racc, dgamma, dbeta, training)
        if dgamma is not None:
            self.grads.ready(spec.module.weight)
        if dbeta is not None:
            self.grads.ready(spec.module.bias)

    def gap(self, x):
        """AdaptiveAvgPool2d(1): [N,H,W,C] -> fp32 [N,1,1,C]."""
        self._materialize(x)
        n, h, w, c = x.shape
        x.consumed += 1
        m = self.be.gap_fwd(x.t)
        out = Act(m.view(n, 1, 1, c))
        if self.record:
            self.tape.append(lambda: self._gap_bwd(x, out))
        return out

    def _gap_bwd(self, x, out):
        if out.grad is None or not x.needs_grad:
            return
        n, h, w, c = x.shape
        dx, acc = self.grad_target(x)
        self.be.gap_bwd(out.grad.reshape(n, c), dx, acc)

    def broadcast(self, src, out):
        """F.interpolate(1x1 -> HxW, bilinear, align_corners=True) == broadcast (DX:450)."""
        n, _, _, c = src.shape
        src.consumed += 1
        self.be.broadcast_hw(src.t.reshape(n, c), out.t)
        if self.record:
            self.tape.append(lambda: self._broadcast_bwd(src, out))
        return out

    def _broadcast_bwd(self, src, out):
        g = out.grad
        if g is None or not src.needs_grad:
            return
        n, _, _, c = src.shape
        if src.grad is not None:
            raise RuntimeError("broadcast source gradient must have a single writer")
        src.grad = self.be.reduce_hw(g).view(n, 1, 1, c)

    def bilinear(self, x, ho, wo, out=None, out_dtype=None):
        """F.interpolate(x, size=(ho, wo), mode='bilinear', align_corners=True) (DX:327-331); `out` may be a concat slice."""
        self._materialize(x)
        n, _, _, c = x.shape
        x.consumed += 1
        if out is None:
            out = self.new_act(n, ho, wo, c, out_dtype or x.t.dtype)
        elif out.shape != (n, ho, wo, c):
            raise RuntimeError("bilinear: output buffer %s does not match %s" % (out.shape, (n, ho, wo, c)))
        self.be.bilinear_fwd(x.t, out.t)
        if self.record:
            self.tape.append(lambda: self._bilinear_bwd(x, out))
        return out

    def _bilinear_bwd(self, x, out):
        g = out.grad
        if g is None or not x.needs_grad:
            return
        dx, acc = self.grad_target(x)
        self.be.bilinear_bwd(g, dx, acc)

    def fork(self, i):
        """Context manager: the operations emitted inside form forward branch i, independent of the other branches (see
        CudaBackend.fork).  The backward tape is unaffected: it replays every closure in reverse order on the main stream."""
        f = getattr(self.be, "fork", None)
        if f is None:
            import contextlib
            return contextlib.nullcontext()
        return f(i)

    def join(self):
        j = getattr(self.be, "join_forks", None)
        if j is not None:
            j()

    # ---- execution ----------------------------------------------------------------------------------------
    def finish_forward(self):
        if self.bn_trained:
            self.be.increment_counters([m.num_batches_tracked for m in self.bn_trained])
            self.bn_trained = []

    def backward(self):
        for fn in reversed(self.tape):
            fn()
        self.tape = []


# ------------------------------------------------------------------------------------------------------------
# autograd bridge
# ------------------------------------------------------------------------------------------------------------
class _PlanFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, precision, n_in, need_grad, *tensors):
        inputs, params = tensors[:n_in], tensors[n_in:]
        device = inputs[0].device
        be = _make_backend(precision, device)
        if hasattr(be, "deterministic"):
            be.deterministic = deterministic()
        if getattr(module, "_dc_grad_sync", None) is not None and hasattr(be, "onepass_bwd"):
            be.onepass_bwd = False       # NCCL kernels share the SMs during backward: no inter-block barriers there
        need_grad = bool(need_grad)
        grads = None
        if need_grad:
            grads = getattr(module, "_dc_gradstore", None)
            if grads is None or not grads.matches(params, device):
                grads = GradStore(params, device)
                module._dc_gradstore = grads
        eng = Engine(be, need_grad, grads)
        xin = [eng.from_nchw(x.detach(), needs_grad=bool(x.requires_grad and need_grad)) for x in inputs]
        outs = module._emit_root(eng, *xin)          # list of (Act, valid channels)
        eng.finish_forward()
        results = tuple(be.to_nchw_f32(act.t, c) for act, c in outs)
        module._dc_last_launches = be.launches
        if need_grad:
            ctx.eng, ctx.outs, ctx.xin, ctx.module = eng, outs, xin, module
            ctx.in_dtypes = [x.dtype for x in inputs]
        else:
            ctx.eng = None
        return results

    @staticmethod
    def backward(ctx, *gouts):
        eng = ctx.eng
        if eng is None:
            raise RuntimeError("backward through a deepcam_b200 module that ran without gradient recording")
        be = eng.be
        be.launches = 0
        flat = eng.grads.begin_backward()
        be.fill_zero_flat(flat)
        sync = getattr(ctx.module, "_dc_grad_sync", None)
        if sync is not None:
            sync.begin(eng.grads)
        for (act, c), g in zip(ctx.outs, gouts):
            if g is not None:
                act.grad = be.from_nchw(g.contiguous(), act.t.shape[3] if act.t.shape[3] != c else None)
        eng.backward()
        if sync is not None:
            sync.finish(eng.grads)
        dxs = []
        for xa, dt in zip(ctx.xin, ctx.in_dtypes):
            if xa.needs_grad and xa.grad is not None:
                dxs.append(be.to_nchw_f32(xa.grad, xa.t.shape[3]).to(dt))
            else:
                dxs.append(None)
        ctx.module._dc_last_launches_bwd = be.launches
        grads = eng.grads.take_views()
        ctx.eng = ctx.outs = ctx.xin = None
        return (None, None, None, None) + tuple(dxs) + grads


# ------------------------------------------------------------------------------------------------------------
# CUDA-graph plans: the whole forward (and backward) of one module call is captured once per configuration and
# replayed afterwards, so the ~1000 kernel launches of a step cost one cudaGraphLaunch each way on the host.
# ------------------------------------------------------------------------------------------------------------
_FWD_ARENA_BYTES = 8 << 20
_BWD_ARENA_BYTES = 192 << 20


def graphs_enabled():
    return os.environ.get("DEEPCAM_B200_GRAPHS", "1") not in ("0", "false", "False", "")


class _capture:
    """torch.cuda.graph with the cyclic garbage collector paused: a collection that happens to run in the middle of a
    capture can finalize an older plan (CUDA graphs + their private memory pool -> cudaFree), which is not permitted on a
    capturing thread and invalidates the capture (seen as 'operation failed due to a previous error during capture')."""

    def __init__(self, graph, pool):
        self.ctx = torch.cuda.graph(graph, pool=pool, capture_error_mode="thread_local")

    def __enter__(self):
        import gc
        self.gc_was_enabled = gc.isenabled()
        gc.collect()
        gc.disable()
        try:
            return self.ctx.__enter__()
        except BaseException:
            if self.gc_was_enabled:
                gc.enable()
            raise

    def __exit__(self, *exc):
        import gc
        try:
            return self.ctx.__exit__(*exc)
        finally:
            if self.gc_was_enabled:
                gc.enable()


class _GraphPlan:
    """Static buffers + captured graphs of one (module, input shapes, mode) configuration.

    Activations, gradient buffers and BatchNorm scratch live in the plan's private memory pool at fixed addresses;
    the user's input is copied into a static NHWC buffer before the forward graph is replayed and the outputs are
    converted into fresh NCHW tensors after it.  One forward may be in flight per plan: backward must consume the
    most recent forward (checked through a generation counter)."""

    def __init__(self, module, precision, device, inputs, params, need_grad):
        from . import _lib
        self._lib = _lib
        self.module = module
        self.device = device
        self.need_grad = need_grad
        self.be = CudaBackend(_PRECISIONS[precision], device)
        self.be.graph_mode = True
        self.be.deterministic = deterministic()
        if getattr(module, "_dc_grad_sync", None) is not None:
            self.be.onepass_bwd = False  # NCCL kernels share the SMs during backward: no inter-block barriers there
        if need_grad and os.environ.get("DEEPCAM_B200_SIDE_BRANCH", "1") not in ("0", "false", ""):
            self.be.side_stream = torch.cuda.Stream(device=device)
        self.pool = torch.cuda.graph_pool_handle()
        self.params = list(params)
        self.guard = _pointer_guard(module, self.params)
        self.grads = GradStore(self.params, device) if need_grad else None
        self.eng = Engine(self.be, need_grad, self.grads)
        self.xin = []
        for x in inputs:
            n, c, h, w = x.shape
            self.xin.append(Act(self.be.empty(n, h, w, c), needs_grad=bool(x.requires_grad and need_grad)))
        self.in_dtypes = [x.dtype for x in inputs]
        self.generation = 0
        self.fwd_graph = None
        self.bwd_segments = None
        self.fwd_kernels = self.bwd_kernels = 0
        self.fwd_zero = self.bwd_zero = None
        self.outs = None

    # ---- forward ------------------------------------------------------------------------------------------
    def _load_inputs(self, inputs):
        for x, act in zip(inputs, self.xin):
            src = x.detach()
            if src.dtype not in (torch.float32, torch.bfloat16):
                src = src.float()
            ops_copy_view(src.permute(0, 2, 3, 1), act.t)

    def forward(self, inputs):
        self._load_inputs(inputs)
        if self.fwd_graph is None:
            from . import ops
            from .backend import pack_jobs_of
            jobs = pack_jobs_of(self.module)                 # all kernel-layout weight copies (filled by the eager call)
            self.pack_table = ops.build_pack_table(jobs, self.device) if jobs else None
            g = torch.cuda.CUDAGraph()
            l0 = self._lib.launch_count
            self.be.begin_arena(_FWD_ARENA_BYTES)
            with _capture(g, self.pool):
                if self.pack_table is not None:
                    ops.pack_weights_multi(*self.pack_table)     # one launch re-packs every weight each step
                    self.be.launches += 1
                    self.be.pack_mark = self.be.launches         # the first kernel after it must not touch weights early
                self.outs = self.module._emit_root(self.eng, *self.xin)
                self.eng.finish_forward()
            self.fwd_zero = self.be.end_arena()
            self.fwd_kernels = self._lib.launch_count - l0
            self.fwd_graph = g
        else:
            self._lib.launch_count += self.fwd_kernels
        if self.fwd_zero is not None:
            self.be.fill_zero_flat(self.fwd_zero)            # BatchNorm workspaces of the whole forward: one memset
        self.fwd_graph.replay()
        self.generation += 1
        self.module._dc_last_launches = self.fwd_kernels + len(self.xin) + len(self.outs)
        return tuple(self.be.to_nchw_f32(act.t, c) for act, c in self.outs)

    # ---- backward -----------------------------------------------------------------------------------------
    def backward(self, gouts):
        be, eng, grads = self.be, self.eng, self.grads
        sync = getattr(self.module, "_dc_grad_sync", None)
        first = self.bwd_segments is None
        if grads.flat is not None:
            # gradient accumulation (a second backward without zero_grad(set_to_none=True)): the .grad tensors the caller still
            # holds are views of the plan's static flat buffer, which the replay below zeroes and overwrites.  Move the held
            # gradients out first; autograd then adds this backward's views to them (g_old + g_new, as in the eager engine).
            lo, hi = grads.flat.data_ptr(), grads.flat.data_ptr() + grads.flat.numel() * 4
            for p in self.params:
                g = p.grad
                if g is not None and lo <= g.data_ptr() < hi:
                    p.grad = g.clone()
        if first:
            flat = grads.begin_backward()
            for (act, c), g in zip(self.outs, gouts):
                if g is not None:
                    n, h, w, cp = act.t.shape
                    act.grad = be.empty(n, h, w, cp)          # backend dtype, like from_nchw in the eager path
            self.gout_used = [g is not None for g in gouts]
        elif [g is not None for g in gouts] != self.gout_used:
            raise RuntimeError("deepcam_b200: the set of outputs receiving gradients changed between backward calls of a "
                               "captured plan; set DEEPCAM_B200_GRAPHS=0 for this usage")
        for (act, c), g in zip(self.outs, gouts):
            if g is not None:
                ops_copy_view(g.detach().permute(0, 2, 3, 1), act.grad)        # pad channels are zero-filled
        if sync is not None:
            sync.begin(grads)
        if not first and self.bwd_zero is not None:
            be.fill_zero_flat(self.bwd_zero)                 # BatchNorm / weight-gradient scratch of the whole backward
        if first:
            tape = list(reversed(eng.tape))
            eng.tape = []
            segments = []
            l0 = self._lib.launch_count
            be.begin_arena(_BWD_ARENA_BYTES)
            be.fill_zero_flat(be.arena)                      # (first call only: the used prefix is not known yet)
            i = 0
            while i < len(tape) or not segments:
                g = torch.cuda.CUDAGraph()
                if sync is not None:
                    sync.deferred = []
                with _capture(g, self.pool):
                    if i == 0:
                        be.fill_zero_flat(grads.flat)
                    while i < len(tape):
                        tape[i]()
                        i += 1
                        if sync is not None and sync.deferred:
                            break
                    be.join_side()                          # the weight-gradient branch rejoins before the segment ends
                buckets = []
                if sync is not None:
                    buckets, sync.deferred = sync.deferred, None
                segments.append((g, buckets))
                g.replay()
                for b in buckets:
                    sync.launch_bucket(b)
            self.bwd_zero = be.end_arena()
            self.bwd_kernels = self._lib.launch_count - l0
            self.bwd_segments = segments
        else:
            self._lib.launch_count += self.bwd_kernels
            for g, buckets in self.bwd_segments:
                g.replay()
                for b in buckets:
                    sync.launch_bucket(b)
        if sync is not None:
            sync.finish(grads)
        dxs = []
        for xa, dt in zip(self.xin, self.in_dtypes):
            if xa.needs_grad and xa.grad is not None:
                dxs.append(be.to_nchw_f32(xa.grad, xa.t.shape[3]).to(dt))
            else:
                dxs.append(None)
        self.module._dc_last_launches_bwd = self.bwd_kernels
        # parameter gradients: fresh views of the static flat buffer (held gradients were moved out of it above)
        return dxs, grads.take_views()


def ops_copy_view(src, dst):
    from . import ops
    ops.copy_view(src, dst)


def _bn_modules(module):
    bns = module.__dict__.get("_dc_bn_list")
    if bns is None:
        bns = [m for m in module.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]
        module.__dict__["_dc_bn_list"] = bns
    return bns


def _pointer_guard(module, params):
    """Addresses baked into a captured graph: parameters and buffers must stay where they are."""
    return tuple(p.data_ptr() for p in params) + tuple(b.data_ptr() for b in module.buffers())


class _GraphFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, n_in, *tensors):
        results = plan.forward(tensors[:n_in])
        ctx.plan = plan
        ctx.gen = plan.generation
        return results

    @staticmethod
    def backward(ctx, *gouts):
        plan = ctx.plan
        if not plan.need_grad:
            raise RuntimeError("backward through a deepcam_b200 module that ran without gradient recording")
        if ctx.gen != plan.generation:
            raise RuntimeError("deepcam_b200: backward() must follow the forward() it belongs to when CUDA-graph plans "
                               "are enabled (a newer forward of the same module has overwritten the saved activations); "
                               "set DEEPCAM_B200_GRAPHS=0 to keep several forwards alive")
        dxs, grads = plan.backward(gouts)
        return (None, None) + tuple(dxs) + grads


def _graph_plan(module, precision, inputs, params, need_grad):
    """Returns the captured plan for this configuration, or None while it is still warming up (first call runs
    eagerly: it fills the kernel-attribute / counter-table caches that must not be touched during capture)."""
    device = inputs[0].device
    if device.type != "cuda" or getattr(_state, "factory", None) is not None or not graphs_enabled():
        return None
    key = (precision, need_grad, tuple(tuple(x.shape) + (x.dtype, bool(x.requires_grad)) for x in inputs),
           tuple(m.training for m in _bn_modules(module)), tuple(p.requires_grad for p in params), device.index,
           torch.cuda.current_stream(device).cuda_stream, getattr(module, "_dc_plan_key", None), deterministic())
    plans = module.__dict__.setdefault("_dc_plans", {})
    ent = plans.get(key)
    if ent is None:
        if len(plans) >= 8:
            plans.clear()
        plans[key] = [1, None]
        return None
    if ent[1] is None:
        ent[1] = _GraphPlan(module, precision, device, inputs, params, need_grad)
    elif ent[1].guard != _pointer_guard(module, params):
        plans[key] = [1, None]            # parameters/buffers moved: run eagerly once, then capture again
        return None
    return ent[1]


def run_module(module, inputs, precision=None):
    """Execute `module` (any class of architecture/deeplab_xception.py) on NCHW inputs through the engine.
    Returns the tuple of NCHW fp32 outputs."""
    for x in inputs:
        if not isinstance(x, torch.Tensor) or x.dim() != 4:
            raise ValueError("expected [N, C, H, W] tensors")
    precision = precision or getattr(module, "precision", None) or default_precision()
    if precision not in _PRECISIONS:
        raise ValueError("precision must be one of %s" % sorted(_PRECISIONS))
    params = list(module.parameters())
    need_grad = torch.is_grad_enabled() and (any(x.requires_grad for x in inputs) or any(p.requires_grad for p in params))
    plan = _graph_plan(module, precision, inputs, params, need_grad)
    if plan is not None:
        return _GraphFunction.apply(plan, len(inputs), *inputs, *params)
    return _PlanFunction.apply(module, precision, len(inputs), need_grad, *inputs, *params)
