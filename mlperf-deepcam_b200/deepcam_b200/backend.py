"""CUDA backend: maps the engine's layer-level operations onto the C-ABI kernels.

There is exactly one product backend and it needs a B200: no CPU path, no cuDNN/cuBLAS path.  (tests/ holds
a torch-CPU interpreter of the same engine operations, used only to check the graph logic without a GPU.)

Layer specs (ConvSpec / DwSpec / BnSpec) reference the fp32 master parameters owned by the nn.Modules of
architecture/deeplab_xception.py; kernel-layout copies (bf16, packed) are caches keyed on the parameter version.
"""
import contextlib
import os

import torch

from . import convdesc, ops
from ._lib import (DC_BN_IDENTITY, DC_BN_MASK_FROM_Y, DC_BN_RELU, DC_BN_RES_WRITE, DC_BN_SUMS_READY, DC_BN_TRAIN,
                   DC_CONV_HALO_PACK, DC_CONV_WEIGHTS_STABLE, DC_PACK_NTK, DC_PACK_NTK_CONVT2, DC_PACK_TKN)


def _round_up(a, b):
    return (a + b - 1) // b * b


class ConvSpec:
    """Dense convolution (nn.Conv2d) or transposed convolution (nn.ConvTranspose2d k3 s2 p1 op1)."""

    def __init__(self, name, weight, bias=None, stride=1, pad=0, dil=1, transposed=False):
        self.name = name
        self.weight = weight            # nn.Parameter: Conv2d [Co,Ci,k,k]; ConvTranspose2d [Ci,Co,k,k]
        self.bias = bias
        self.stride, self.pad, self.dil = stride, pad, dil
        self.transposed = transposed
        if transposed:
            self.ci, self.co = weight.shape[0], weight.shape[1]
        else:
            self.co, self.ci = weight.shape[0], weight.shape[1]
        self.k = weight.shape[2]
        self._cache = {}

    def out_hw(self, h, w):
        if self.transposed:   # output_padding = 1 (DX:352): (h-1)*s - 2p + k + 1
            f = lambda v: (v - 1) * self.stride - 2 * self.pad + self.k + 1
        else:
            f = lambda v: convdesc.conv_out_size(v, self.k, self.stride, self.pad, self.dil)
        return f(h), f(w)


class DwSpec:
    """Depthwise 3x3 of SeparableConv2d_same (DX:58-59) with fixed_padding (DX:45-51)."""

    def __init__(self, name, weight, stride=1, dil=1):
        self.name = name
        self.weight = weight            # [C,1,3,3]
        self.c = weight.shape[0]
        self.stride, self.dil = stride, dil
        self._cache = {}

    def out_hw(self, h, w):
        return (h - 1) // self.stride + 1, (w - 1) // self.stride + 1


class BnSpec:
    def __init__(self, name, module):
        self.name = name
        self.module = module            # nn.BatchNorm2d holding weight/bias/running stats

    @property
    def c(self):
        return self.module.num_features


_COUNTER_TABLES = {}      # tuple of counter addresses -> device table (built once per module instance)


def _cached(spec, key, param, job, frozen=False):
    """Kernel-layout copy of a parameter.  The buffer is allocated once and re-packed IN PLACE whenever the master
    parameter changed, so its address stays valid for captured CUDA graphs.  `job` = (K, N, taps, src_k_first, layout,
    K_pad, N_pad, dtype) of ops.pack_weight.  `frozen` (graph capture): an existing buffer is returned as it is,
    because the plan re-packs every registered buffer in one launch at the start of its forward graph."""
    ver = (param._version, param.data_ptr())
    ent = spec._cache.get(key)
    if ent is None or ent[1].device != param.device:
        ent = [ver, ops.pack_weight(param.detach(), *job), job, param]
        spec._cache[key] = ent
    elif not frozen and ent[0] != ver:
        ops.pack_weight(param.detach(), *job, out=ent[1])
        ent[0] = ver
    return ent[1]


def pack_jobs_of(module):
    """Every kernel-layout weight copy that exists for the layers under `module` (filled by an eager run)."""
    jobs = []
    for m in module.modules():
        for spec in m.__dict__.get("_dc_specs", {}).values():
            for ent in getattr(spec, "_cache", {}).values():
                K, N, taps, skf, layout, K_pad, N_pad, _dt = ent[2]
                jobs.append((ent[3].detach(), ent[1], K, N, taps, skf, layout, K_pad, N_pad))
    return jobs


class CudaBackend:
    name = "cuda"

    def __init__(self, dtype=torch.bfloat16, device=None, use_tc=None):
        if not torch.cuda.is_available():
            raise RuntimeError("deepcam_b200: CUDA device required (the kernels are sm_100a only; no CPU fallback)")
        self.dtype = dtype
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        if use_tc is None:
            use_tc = dtype == torch.bfloat16 and ops._lib.load().dc_device_supports_tcgen05() == 1
        self.use_tc = bool(use_tc)
        self.launches = 0
        # bit-identical gradients from run to run (engine.deterministic()): two-stage weight-gradient reductions through a
        # workspace, single-writer pooling reduction
        self.deterministic = False
        self.graph_mode = False       # True while the owning plan captures CUDA graphs (engine._GraphPlan)
        self.arena = None
        self.arena_used = 0
        # one-launch BatchNorm for on-chip-sized tensors (inter-block barrier; needs all blocks co-resident).  The
        # backward variant is switched off by the data-parallel wrapper: NCCL kernels share the SMs during backward.
        self.onepass = os.environ.get("DEEPCAM_B200_BN_ONEPASS", "1") not in ("0", "false", "")
        self.onepass_bwd = os.environ.get("DEEPCAM_B200_BN_ONEPASS_BWD", "1") not in ("0", "false", "")
        # BatchNorm batch sums out of the producing GEMM's epilogue (dc_conv_gemm_tc_bnstats): no statistics pass at all
        self.fuse_bn_stats = os.environ.get("DEEPCAM_B200_FUSE_BN_STATS", "1") not in ("0", "false", "")
        # few-channel stride-2 ConvTranspose2d forward as ONE 2x2-tap contraction over all four output parities
        self.fuse_convT = os.environ.get("DEEPCAM_B200_FUSE_CONVT", "1") not in ("0", "false", "")
        # BatchNorm backward reduction (+ ReLU mask) inside the depthwise backward-data kernel that produces the gradient
        self.fuse_bn_bwd = os.environ.get("DEEPCAM_B200_FUSE_BN_BWD", "0") not in ("0", "false", "")
        self.fuse_bn_bwd_max_bytes = int(os.environ.get("DEEPCAM_B200_FUSE_BN_BWD_MAX_BYTES", str(24 << 20)))
        self.bn_bwd_split = os.environ.get("DEEPCAM_B200_BN_BWD_SPLIT", "1") not in ("0", "false", "")
        self.bn_bwd_split_max_bytes = int(os.environ.get("DEEPCAM_B200_BN_BWD_SPLIT_MAX_BYTES", str(1 << 40)))
        self.fuse_bn_bwd_res = os.environ.get("DEEPCAM_B200_FUSE_BN_BWD_RES", "1") not in ("0", "false", "")
        # BatchNorm(+ReLU) applied while the following depthwise kernel loads its tile (dc_dw_fwd_bn): no bn_apply launch
        # Opt-in: measured on B200 (round 2, tools/kbench.py, 2x48x72x728): 16.1 us fused vs 8.6 us dw_fwd + 5.3 us bn_apply - the
        # in-place activation pass over the staged tile and its second block barrier cost more than the launch they save
        self.fuse_bn_dw = os.environ.get("DEEPCAM_B200_FUSE_BN_DW", "0") not in ("0", "false", "")
        # eval mode without gradient recording: BatchNorm (+ReLU) folded into the producing GEMM's epilogue
        self.fold_bn_eval = os.environ.get("DEEPCAM_B200_FOLD_BN_EVAL", "1") not in ("0", "false", "")
        # tcgen05 GEMMs fetch their first weight tiles before griddepcontrol.wait when the weights are known to be older than
        # the preceding kernel (captured plans only, see _weights_stable)
        self.early_weights = os.environ.get("DEEPCAM_B200_EARLY_WEIGHTS", "1") not in ("0", "false", "")
        self.pack_mark = 0            # value of self.launches right after the most recent weight-pack launch
        self.fork_branches = os.environ.get("DEEPCAM_B200_FORK_ASPP", "1") not in ("0", "false", "")
        self._fork_streams, self._fork_dirty = [], set()
        self.side_stream = None       # set by a graph plan: weight-gradient kernels run on a parallel graph branch
        self._side_dirty = False

    # ---- memory -----------------------------------------------------------------------------------
    def empty(self, n, h, w, c, dtype=None):
        return torch.empty((n, h, w, c), dtype=dtype or self.dtype, device=self.device)

    def zeros(self, n, h, w, c, dtype=None):
        t = self.empty(n, h, w, c, dtype)
        ops.fill_zero(t)
        self.launches += 1
        return t

    def scratch(self, numel, dtype, zero=False):
        if zero and self.arena is not None:
            # graph capture: zero-initialised scratch comes out of one arena that the plan clears with a single
            # memset before every replay (instead of one memset node per BatchNorm workspace / gradient scratch)
            nbytes = numel * torch.empty((), dtype=dtype).element_size()
            off = (self.arena_used + 255) // 256 * 256
            if off + nbytes <= self.arena.numel():
                self.arena_used = off + nbytes
                return self.arena[off:off + nbytes].view(dtype)
        t = torch.empty(numel, dtype=dtype, device=self.device)
        if zero:
            ops.fill_zero(t)
            self.launches += 1
        return t

    @contextlib.contextmanager
    def side_branch(self):
        """Weight-gradient kernels only feed the optimizer, so inside a captured backward they run on a second stream
        (a parallel branch of the CUDA graph) and overlap with the latency-bound main chain of data-gradient
        kernels.  Everything they read was produced on the main stream before the fork; buffers stay alive for the
        whole capture and zero-initialised scratch comes from the arena, so there is no reuse hazard."""
        if self.side_stream is None:
            yield
            return
        main = torch.cuda.current_stream(self.device)
        ev = torch.cuda.Event()
        ev.record(main)
        self.side_stream.wait_event(ev)
        self._side_dirty = True
        with torch.cuda.stream(self.side_stream):
            yield

    @contextlib.contextmanager
    def fork(self, i):
        """Independent FORWARD branches (the four ASPP branches, DX:443-446, read the same feature map and write disjoint channel
        slices of the concat buffer): inside a captured plan branch i runs on its own stream = a parallel branch of the CUDA graph,
        so the 54-CTA atrous GEMMs of three branches share the 148 SMs instead of running one after the other.  join_forks()
        brings them back before the first consumer of the concat buffer.  Outside a capture this is a no-op."""
        if not self.graph_mode or not self.fork_branches or not torch.cuda.is_current_stream_capturing():
            yield
            return
        while len(self._fork_streams) <= i:
            self._fork_streams.append(torch.cuda.Stream(device=self.device))
        main = torch.cuda.current_stream(self.device)
        ev = torch.cuda.Event()
        ev.record(main)
        st = self._fork_streams[i]
        st.wait_event(ev)
        self._fork_dirty.add(i)
        with torch.cuda.stream(st):
            yield

    def join_forks(self):
        main = torch.cuda.current_stream(self.device)
        for i in sorted(self._fork_dirty):
            ev = torch.cuda.Event()
            ev.record(self._fork_streams[i])
            main.wait_event(ev)
        self._fork_dirty.clear()

    def join_side(self):
        if self.side_stream is not None and self._side_dirty:
            ev = torch.cuda.Event()
            ev.record(self.side_stream)
            torch.cuda.current_stream(self.device).wait_event(ev)
            self._side_dirty = False

    def begin_arena(self, nbytes):
        self.arena = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.arena_used = 0
        return self.arena

    def end_arena(self):
        """Returns the used prefix of the arena (to be zeroed before every replay) and detaches it."""
        used = self.arena[:(self.arena_used + 255) // 256 * 256] if self.arena_used else None
        self.arena = None
        return used

    def fill_zero_flat(self, flat):
        ops.fill_zero(flat)
        self.launches += 1

    def increment_counters(self, counters):
        """num_batches_tracked += 1 for every BatchNorm that ran in training mode: one launch for all of them."""
        key = tuple(c.data_ptr() for c in counters)
        table = _COUNTER_TABLES.get(key)
        if table is None:
            table = torch.tensor(key, dtype=torch.int64, device=self.device)
            if len(_COUNTER_TABLES) > 64:
                _COUNTER_TABLES.clear()
            _COUNTER_TABLES[key] = table
        ops.i64_increment_many(table, len(key))
        self.launches += 1

    # ---- layout -------------------------------------------------------------------------------------
    def from_nchw(self, x_nchw, c_pad=None):
        n, c, h, w = x_nchw.shape
        out = self.empty(n, h, w, c_pad or c)
        src = x_nchw if x_nchw.dtype in (torch.float32, torch.bfloat16) else x_nchw.float()
        ops.copy_view(src.permute(0, 2, 3, 1), out)
        self.launches += 1
        return out

    def to_nchw_f32(self, act, c):
        n, h, w, _ = act.shape
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=self.device)
        ops.copy_view(act[..., :c] if act.shape[3] != c else act, out.permute(0, 2, 3, 1))
        self.launches += 1
        return out

    # ---- helpers --------------------------------------------------------------------------------------
    def _tc_ok(self, gathered, other_c):
        """tcgen05 path eligibility for a contraction whose gathered operand is `gathered`."""
        if not self.use_tc or gathered.dtype != torch.bfloat16:
            return False
        c = gathered.shape[3]
        if c % 8 or c < 8 or other_c < 1:
            return False
        if gathered.stride(3) != 1 or any(s % 8 for s in gathered.stride()[:3]) or gathered.data_ptr() % 16:
            return False
        return gathered.shape[0] * gathered.shape[1] * gathered.shape[2] >= 128

    def _halo(self, taps, stride, wtaps, x, out):
        """DC_CONV_HALO_PACK when the tcgen05 kernel will run this gather in halo mode AND the gathered channel count is not a
        multiple of 64 (then the weights must be packed K-dense, see _packed(k_dense=True)); 0 otherwise."""
        if x.shape[3] % 64 == 0:
            return 0                                    # dense and 64-padded packs coincide: the C side decides alone
        desc = ops.make_desc(taps, (stride, stride), False, wtaps)
        return DC_CONV_HALO_PACK if ops.conv_halo_ok(desc, x, out) else 0

    def _packed(self, spec, role, impl, x, n_pad=None, k_dense=False):
        """role: 'fprop' (k = op input channels) or 'dgrad' (k = op output channels); x = gathered operand.
        k_dense: halo-mode layout [N][slice][C] with C = the gathered operand's channels instead of a multiple of 64."""
        w = spec.weight
        taps = spec.k * spec.k
        if role == "fprop":
            K, N = spec.ci, spec.co
            src_k_first = spec.transposed          # ConvT weight is [ci][co][taps] = [k][n]
        else:
            K, N = spec.co, spec.ci
            src_k_first = not spec.transposed      # Conv weight is [co][ci][taps] = [k][n]
        if impl == "tc":
            layout, K_pad, N_pad, dt = DC_PACK_NTK, (x.shape[3] if k_dense else _round_up(K, 64)), max(N, n_pad or 0), torch.bfloat16
        else:
            layout, K_pad, N_pad, dt = DC_PACK_TKN, x.shape[3], _round_up(N, 4), x.dtype

        c0 = ops._lib.launch_count
        buf = _cached(spec, (role, impl, dt, K_pad, N_pad), w, (K, N, taps, src_k_first, layout, K_pad, N_pad, dt),
                      frozen=self.graph_mode)
        if ops._lib.launch_count != c0:
            self.pack_mark = self.launches          # a pack kernel was just launched: the next kernel follows it directly
        return buf

    def _weights_stable(self):
        """DC_CONV_WEIGHTS_STABLE for the next GEMM: inside a captured plan every packed weight was refreshed by the ONE pack
        launch at the start of the forward graph, so any kernel with at least one other launch between it and that pack may
        fetch its weights before griddepcontrol.wait (the early portion of a kernel only overlaps its immediate predecessor)."""
        return DC_CONV_WEIGHTS_STABLE if (self.graph_mode and self.early_weights and self.launches > self.pack_mark) else 0

    def _gemm(self, taps, stride, accumulate, wtaps, x, w, bias, out, impl, bn_sums=None, out_split=None, flop_scale=1.0, flags=0):
        desc = ops.make_desc(taps, (stride, stride), accumulate, wtaps, out_split,
                             flags=(flags | self._weights_stable()) if impl == "tc" else 0)
        ops.conv_gemm(desc, x, w, bias, out, impl, bn_sums, flop_scale)
        self.launches += 1

    # ---- dense convolution ---------------------------------------------------------------------------------
    def conv_fwd(self, x, spec, out, want_bn_sums=False):
        """out <- conv(x) (+ bias).  `out` is a preallocated logical-NHWC tensor (possibly a channel slice).
        want_bn_sums: the caller will batch-normalise `out` in training mode; when the tcgen05 kernel runs, its epilogue
        also accumulates the per-channel batch sums and the (zeroed) BatchNorm workspace holding them is returned
        (pass it to bn_fwd as `ready_sums`); otherwise None is returned and bn_fwd computes the statistics itself."""
        impl = "tc" if self._tc_ok(x, spec.co) else "simt"
        if spec.transposed and impl == "tc" and not want_bn_sums and self._convT_fusable(spec, out):
            return self._convT_fused_fwd(x, spec, out)
        kk = spec.k * spec.k
        halo = 0
        if impl == "tc" and not spec.transposed and kk > 1:
            halo = self._halo(convdesc.conv_fprop_taps(spec.k, spec.pad, spec.dil), spec.stride, kk, x, out)
        w = self._packed(spec, "fprop", impl, x, n_pad=out.shape[3], k_dense=bool(halo))
        bias = spec.bias.detach() if spec.bias is not None else None
        sums = None
        if want_bn_sums and impl == "tc" and self.fuse_bn_stats and out.dtype == torch.bfloat16:
            sums = self.scratch(ops.bn_ws_elems(out.shape[3]), torch.float64, zero=True)
        if not spec.transposed:
            self._gemm(convdesc.conv_fprop_taps(spec.k, spec.pad, spec.dil), spec.stride, False, kk, x, w, bias, out, impl, sums,
                       flags=halo)
        else:
            s = spec.stride
            for ph in range(s):
                for pw in range(s):
                    taps = convdesc.convT_fprop_taps(spec.k, s, spec.pad, ph, pw)
                    self._gemm(taps, 1, False, kk, x, w, bias, out[:, ph::s, pw::s, :], impl, sums)
        return sums

    def conv_bn_eval_fwd(self, x, spec, bnspec, relu, out):
        """out <- [relu](bn_eval(conv(x))) with the eval-mode BatchNorm folded into the tcgen05 epilogue (no pre-BatchNorm tensor,
        no bn_apply launch).  Returns False - nothing launched - when this layer cannot take that path (SIMT layers, fp32
        mode, output layouts that need the generic epilogue); the caller then runs conv_fwd + bn_fwd."""
        m = bnspec.module
        if (not self.fold_bn_eval or m.running_mean is None or m.weight is None or m.bias is None or out.dtype != torch.bfloat16
                or not self._tc_ok(x, spec.co)):
            return False
        kk = spec.k * spec.k
        halo = 0
        if not spec.transposed and kk > 1:
            halo = self._halo(convdesc.conv_fprop_taps(spec.k, spec.pad, spec.dil), spec.stride, kk, x, out)
        w = self._packed(spec, "fprop", "tc", x, n_pad=out.shape[3], k_dense=bool(halo))
        bias = spec.bias.detach() if spec.bias is not None else None
        args = (m.weight.detach(), m.bias.detach(), m.running_mean, m.running_var, m.eps, relu)
        if not spec.transposed:
            desc = ops.make_desc(convdesc.conv_fprop_taps(spec.k, spec.pad, spec.dil), (spec.stride, spec.stride), False, kk,
                                 flags=self._weights_stable() | halo)
            if not ops.conv_gemm_bn_eval(desc, x, w, bias, out, *args):
                return False
            self.launches += 1
            return True
        s = spec.stride
        for ph in range(s):
            for pw in range(s):
                desc = ops.make_desc(convdesc.convT_fprop_taps(spec.k, s, spec.pad, ph, pw), (1, 1), False, kk,
                                     flags=self._weights_stable())
                if not ops.conv_gemm_bn_eval(desc, x, w, bias, out[:, ph::s, pw::s, :], *args):
                    if ph or pw:
                        raise RuntimeError("deepcam_b200: transposed-convolution parity classes disagree on the epilogue path")
                    return False
                self.launches += 1
        return True

    def _convT_fusable(self, spec, out):
        """Few-channel stride-2 transposed convolution (last_deconv, DX:374) into a dense output: the four parity launches
        each re-read the whole input; one 2x2-tap contraction with N = 4 * channels reads it once per tap."""
        g = out.shape[3]
        return (self.fuse_convT and spec.k == 3 and spec.stride == 2 and spec.pad == 1 and spec.bias is None and
                g % 4 == 0 and 4 * g <= 64 and out.stride(3) == 1 and out.stride(2) == g and
                out.shape[1] % 2 == 0 and out.shape[2] % 2 == 0)

    def _convT_fused_fwd(self, x, spec, out):
        g = out.shape[3]
        job = (spec.ci, spec.co, 4, True, DC_PACK_NTK_CONVT2, _round_up(spec.ci, 64), 4 * g, torch.bfloat16)
        w = _cached(spec, ("fprop_convT2", g), spec.weight, job, frozen=self.graph_mode)
        n, h2, w2, _ = out.shape
        # output pixel (2i+a, 2j+b): channels (b, co) of row 2i+a are contiguous, row parity a is the second segment
        view = out.as_strided((n, h2 // 2, w2 // 2, 4 * g), (out.stride(0), 2 * out.stride(1), 2 * out.stride(2), 1),
                              out.storage_offset())
        self._gemm(convdesc.convT_fused_fprop_taps(), 1, False, 4, x, w, None, view, "tc",
                   out_split=(2 * g, out.stride(1)), flop_scale=9.0 / 16.0)      # 9 of the 16 (class, tap) blocks are non-zero
        return None

    def conv_bwd_data(self, dy, spec, dx, accumulate):
        """dx (+)= conv^T(dy)."""
        impl = "tc" if self._tc_ok(dy, spec.ci) else "simt"
        kk = spec.k * spec.k
        halo = 0
        if impl == "tc" and spec.transposed:
            halo = self._halo(convdesc.convT_dgrad_taps(spec.k, spec.pad), spec.stride, kk, dy, dx)
        w = self._packed(spec, "dgrad", impl, dy, k_dense=bool(halo))
        if spec.transposed:
            self._gemm(convdesc.convT_dgrad_taps(spec.k, spec.pad), spec.stride, accumulate, kk, dy, w, None, dx, impl, flags=halo)
        elif spec.stride == 1:
            self._gemm(convdesc.conv_dgrad_taps(spec.k, spec.pad, spec.dil), 1, accumulate, kk, dy, w, None, dx, impl)
        else:
            # strided Conv2d: the input gradient of parity class (ph, pw) is a stride-1 gather of dy
            s = spec.stride
            if not accumulate:
                if dx.is_contiguous():
                    ops.fill_zero(dx)
                else:
                    raise NotImplementedError("strided dgrad into a non-contiguous gradient needs accumulate=True")
                self.launches += 1
            for ph in range(s):
                for pw in range(s):
                    taps = []
                    for kh in range(spec.k):
                        if (ph + spec.pad - kh * spec.dil) % s:
                            continue
                        for kw in range(spec.k):
                            if (pw + spec.pad - kw * spec.dil) % s:
                                continue
                            taps.append(((ph + spec.pad - kh * spec.dil) // s, (pw + spec.pad - kw * spec.dil) // s,
                                         kh * spec.k + kw))
                    sub = dx[:, ph::s, pw::s, :]
                    if taps and sub.shape[1] > 0 and sub.shape[2] > 0:
                        self._gemm(taps, 1, True, kk, dy, w, None, sub, impl)
        return dx

    def conv_bwd_weight(self, x, dy, spec, wgrad, bgrad=None):
        """wgrad <- dL/dW in the parameter's own layout (fp32, contiguous view of the flat gradient buffer).
        The destination must be zero on entry when it is written in place (1x1 Conv2d)."""
        kk = spec.k * spec.k
        if spec.transposed:
            gathered, enumerated = dy, x
            taps = convdesc.convT_wgrad_taps(spec.k, spec.pad)
            K, N = spec.co, spec.ci          # G is [tap][ci_in][co_out]
        else:
            gathered, enumerated = x, dy
            taps = convdesc.conv_wgrad_taps(spec.k, spec.pad, spec.dil)
            K, N = spec.ci, spec.co          # G is [tap][co][ci]
        impl = "tc" if (self._tc_ok(gathered, 16) and self._tc_ok(enumerated, 16)) else "simt"
        desc = ops.make_desc(taps, (spec.stride, spec.stride), False, kk)
        ws = None
        if self.deterministic:
            nws = ops.conv_wgrad_ws_elems(desc, gathered, enumerated, impl)
            if nws > 0:
                ws = self.scratch(nws, torch.float32)
                self.launches += 1               # the ordered second stage
        if kk == 1 and enumerated.shape[3] == N:
            ops.conv_wgrad(desc, gathered, enumerated, wgrad, impl, ws=ws)       # [co][ci] is already the parameter layout
            self.launches += 1
        else:
            ks = gathered.shape[3]               # >= K when the gathered operand is channel-padded (3 -> 4 logits)
            G = self.scratch(kk * ks * enumerated.shape[3], torch.float32, zero=True)
            ops.conv_wgrad(desc, gathered, enumerated, G, impl, ws=ws)
            ops.unpack_wgrad(G, K, N, kk, False, wgrad, k_stride=ks)
            self.launches += 2
        if bgrad is not None:
            ws = self.scratch(dy.shape[3], torch.float64, zero=True)     # dy may carry padding channels beyond spec.co
            ops.channel_sum(dy, ws, bgrad)
            self.launches += 2
        return wgrad

    # ---- depthwise ---------------------------------------------------------------------------------------------
    def _dw_packed(self, spec):
        return _cached(spec, ("dw", self.dtype), spec.weight, (1, spec.c, 9, False, DC_PACK_TKN, 1, spec.c, self.dtype),
                       frozen=self.graph_mode)

    def dw_fwd(self, x, spec, out):
        ops.dw_fwd(x, self._dw_packed(spec), spec.stride, spec.dil, out)
        self.launches += 1
        return out

    def bn_dw_fwd(self, y, bnspec, relu, dwspec, act, out, ready_sums):
        """act <- [relu](bn(y)), out <- depthwise(act) in ONE launch (train-mode BatchNorm whose batch sums came out of the
        producing GEMM's epilogue; stride 1, dilation 1).  Returns the BatchNorm workspace (as bn_fwd does), or None when the
        fused kernel is not applicable - nothing has been launched then and the caller runs bn_fwd + dw_fwd."""
        if not self.fuse_bn_dw or ready_sums is None or dwspec.stride != 1 or dwspec.dil != 1 or y.dtype != self.dtype:
            return None
        m = bnspec.module
        n, h, w, c = y.shape
        if n * h * w <= 1 or m.weight is None or m.bias is None:
            return None
        track = m.running_mean is not None
        mom = m.momentum if m.momentum is not None else 0.1
        flags = DC_BN_TRAIN | DC_BN_SUMS_READY | (DC_BN_RELU if relu else 0)
        p = ops.bn_params(m.weight.detach(), m.bias.detach(), m.running_mean if track else None, m.running_var if track else None,
                          ready_sums, n * h * w, mom, m.eps, flags)
        if not ops.dw_fwd_bn(p, y, self._dw_packed(dwspec), act, out):
            return None
        self.launches += 1
        return ready_sums

    def dw_bwd_data(self, dy, spec, dx, accumulate):
        ops.dw_bwd_data(dy, self._dw_packed(spec), spec.stride, spec.dil, dx, accumulate)
        self.launches += 1
        return dx

    def dw_bwd_data_bnred(self, dy, spec, dx, accumulate, y, act, fwd_sums, relu, force=False):
        """dw_bwd_data as the LAST writer of the gradient of a BatchNorm output: also applies the ReLU mask and leaves the
        BatchNorm backward sums in a fresh workspace, which is returned (None: not applicable, nothing was launched)."""
        if (not self.fuse_bn_bwd and not force) or spec.stride != 1 or spec.dil != 1:
            return None
        # measured on B200 (session 4): the fused kernel (24 us + 5.5 us apply on the 10 MB tensors) beats the one-pass barrier
        # kernel after a plain depthwise backward (9.5 + 20.7 us) but loses to the split reduction (9.5 + 16 us, bn_bwd below),
        # so it is opt-in (DEEPCAM_B200_FUSE_BN_BWD=1); on the large tensors its per-block fp64 atomics outweigh the saved pass
        if not force and ((act is not None and not self.fuse_bn_bwd_res) or
                          dx.numel() * dx.element_size() > self.fuse_bn_bwd_max_bytes):
            return None
        rws = self.scratch(ops.bn_ws_elems(dx.shape[3]), torch.float64, zero=True)
        if not ops.dw_bwd_data_bnred(dy, self._dw_packed(spec), dx, accumulate, y, act, fwd_sums, rws, relu):
            return None
        self.launches += 1
        return rws

    def bn_bwd_reduced(self, g, y, spec, sums, rws, dy, dres, res_accumulate, dgamma, dbeta):
        """BatchNorm backward (train mode) for a gradient that dw_bwd_data_bnred already masked and reduced."""
        m = spec.module
        flags = DC_BN_TRAIN | (0 if res_accumulate else DC_BN_RES_WRITE)
        n, h, w, c = g.shape
        p = ops.bn_params(m.weight.detach(), m.bias.detach(), m.running_mean, m.running_var, sums, n * h * w, 0.0, m.eps, flags)
        ops.bn_bwd_apply_reduced(p, g, y, rws, dy, dres, dgamma, dbeta)
        self.launches += 1

    def dw_bwd_weight(self, x, dy, spec, wgrad):
        """wgrad: fp32 [C,1,3,3] view of the flat gradient buffer, ZERO on entry (the engine clears the flat buffer at the
        start of backward): the kernel accumulates straight into the parameter layout, no scratch and no unpack."""
        ws = None
        if self.deterministic:
            ws = self.scratch(ops.dw_bwd_weight_ws_elems(x, dy, spec.stride, spec.dil), torch.float32)
            self.launches += 1                   # the ordered second stage
        ops.dw_bwd_weight(x, dy, spec.stride, spec.dil, wgrad, param_layout=True, ws=ws)
        self.launches += 1
        return wgrad

    # ---- batch norm (+relu, +residual) ------------------------------------------------------------------------------
    def bn_fwd(self, y, spec, relu, residual, out, training, ready_sums=None):
        """out <- [relu](bn(y) [+ residual]).  spec None = identity (pure relu / add).  Returns the saved statistics.
        ready_sums: workspace returned by conv_fwd(want_bn_sums=True) for the same y (batch sums already accumulated)."""
        flags = DC_BN_RELU if relu else 0
        n, h, w, c = y.shape
        if spec is None:
            p = ops.bn_params(None, None, None, None, None, n * h * w, 0.0, 0.0, flags | DC_BN_IDENTITY)
            ops.bn_apply(p, y, residual, out)
            self.launches += 1
            return None
        m = spec.module
        sums = None
        mom = m.momentum if m.momentum is not None else 0.1
        track = m.running_mean is not None
        if training:
            if n * h * w <= 1:
                raise ValueError("Expected more than 1 value per channel when training, got input size %s"
                                 % (torch.Size((n, c, h, w)),))
            flags |= DC_BN_TRAIN
            if ready_sums is not None:
                sums = ready_sums
                flags |= DC_BN_SUMS_READY
            else:
                sums = self.scratch(ops.bn_ws_elems(c), torch.float64, zero=True)    # per-layer BatchNorm workspace
        p = ops.bn_params(m.weight.detach(), m.bias.detach(), m.running_mean if track else None,
                          m.running_var if track else None, sums, n * h * w, mom, m.eps, flags)
        if training and ready_sums is not None:
            ops.bn_apply(p, y, residual, out)      # finalizes the sums itself (coefficients, running statistics), then applies
            self.launches += 1
            return sums
        if training and self.onepass and ops.bn_onepass_ok(c, n * h * w, y.dtype, False):
            ops.bn_fwd_onepass(p, y, residual, out)     # statistics + normalisation in one launch (tensor held on chip)
            self.launches += 1
            return sums
        if training:
            ops.bn_stats(p, y)            # sums + coefficients + running statistics (last block finalizes)
            self.launches += 1
        ops.bn_apply(p, y, residual, out)
        self.launches += 1
        return sums

    def bn_bwd(self, dout, out, y, spec, sums, relu, dy, dres, res_accumulate, dgamma, dbeta, training=True):
        flags = DC_BN_RELU if relu else 0
        if not res_accumulate:
            flags |= DC_BN_RES_WRITE
        n, h, w, c = dout.shape
        if spec is None:
            p = ops.bn_params(None, None, None, None, None, n * h * w, 0.0, 0.0, flags | DC_BN_IDENTITY)
            ops.bn_bwd_apply(p, dout, out, None, None, dy, dres)
            self.launches += 1
            return
        m = spec.module
        if training:
            flags |= DC_BN_TRAIN
        mask_src = out if relu else None
        if relu and training and dres is None:
            # out = relu(y * scale + shift) with the coefficients still in the forward workspace: the kernels recompute the
            # ReLU decision from y (which they read anyway) instead of reading `out`
            flags |= DC_BN_MASK_FROM_Y
            mask_src = None
        p = ops.bn_params(m.weight.detach(), m.bias.detach(), m.running_mean, m.running_var, sums, n * h * w, 0.0, m.eps, flags)
        rws = self.scratch(ops.bn_ws_elems(c), torch.float64, zero=True)
        if training and self.bn_bwd_split and dout.numel() * dout.element_size() <= self.bn_bwd_split_max_bytes:
            # sums-only reduction + apply that finalizes per block in shared memory (no last-block tail, no inter-block
            # barrier): measured ahead of the one-pass kernel on the 10 MB tensors (16.0 vs 20.7 us) and of reduce-with-finalize
            # + coefficient-loading apply on the large ones (13.44 vs 13.52 ms/step)
            pr = ops.bn_params(m.weight.detach(), m.bias.detach(), m.running_mean, m.running_var, sums, n * h * w, 0.0, m.eps,
                               flags | DC_BN_SUMS_READY)
            ops.bn_bwd_reduce(pr, dout, mask_src, y, rws, None, None)
            ops.bn_bwd_apply_finalize(p, dout, mask_src, y, rws, dy, dres, dgamma, dbeta)
            self.launches += 2
            return
        if self.onepass and self.onepass_bwd and ops.bn_onepass_ok(c, n * h * w, dout.dtype, True):
            ops.bn_bwd_onepass(p, dout, mask_src, y, rws, dy, dres, dgamma, dbeta)
            self.launches += 1
            return
        ops.bn_bwd_reduce(p, dout, mask_src, y, rws, dgamma, dbeta)
        ops.bn_bwd_apply(p, dout, mask_src, y, rws, dy, dres)
        self.launches += 2

    # ---- image pooling branch ------------------------------------------------------------------------------------------
    def gap_fwd(self, x):
        n, _, _, c = x.shape
        out = torch.empty((n, c), dtype=torch.float32, device=self.device)
        ops.set_deterministic(self.deterministic)
        ops.gap_fwd(x, out)
        self.launches += 2
        return out

    def reduce_hw(self, x):
        n, _, _, c = x.shape
        out = torch.empty((n, c), dtype=torch.float32, device=self.device)
        ops.set_deterministic(self.deterministic)
        ops.reduce_hw(x, out)
        self.launches += 2
        return out

    def broadcast_hw(self, src_nc, out):
        ops.broadcast_hw(src_nc, out)
        self.launches += 1
        return out

    def gap_bwd(self, dmean_nc, dx, accumulate):
        ops.gap_bwd(dmean_nc, dx, accumulate)
        self.launches += 1
        return dx

    # ---- bilinear resize (InterpolationUpsampler, DX:327-331) ----------------------------------------------------
    def bilinear_fwd(self, x, out):
        ops.bilinear_fwd(x, out)
        self.launches += 1
        return out

    def bilinear_bwd(self, dout, din, accumulate):
        ops.bilinear_bwd(dout, din, accumulate)
        self.launches += 1
        return din
