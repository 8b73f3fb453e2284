#!/usr/bin/env python
"""bench.py — DeepCAM training-step throughput on B200 (BASELINE.json metric: train samples/s @768x1152x16).

  python bench.py --gpus N --steps K --warmup W            our arm (hand-written sm_100a kernels behind the reference API)
  python bench.py --impl reference --gpus N --steps K ...  reference arm: the CPU oracle port of the reference step

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for the definition of every field.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.abspath(__file__))
_JSON_OUT = sys.stdout
sys.path.insert(0, os.path.join(REPO, "mlperf-deepcam_b200"))

H, W, C_IN, N_CLASSES, LOCAL_BATCH = 768, 1152, 16, 3, 2
METRIC = "train_samples_per_s_768x1152x16"
FWD_BWD_GFLOP_PER_SAMPLE = 1963.976
WORKLOAD = ("configs[1]: DeepLabv3+/Xception (n_input=16, n_classes=3, os=16) fwd + weighted-CE + bwd + Adam, "
            "local batch 2, synthetic 768x1152x16 tiles")


def _peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            with open(self.path) as fh:
                for line in fh:
                    f = [x.strip() for x in line.split(",")]
                    if len(f) < 9 or not f[0].isdigit() or int(f[0]) != self.index:
                        continue
                    try:
                        sm.append(float(f[1])); mx.append(float(f[2]))
                    except ValueError:
                        continue
                    for nm, val in zip(names, f[5:9]):
                        if val.lower().startswith("active"):
                            reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# --------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(REPO, "baseline", "_ref", "deepCam")
CLASS_WEIGHTS = [1.001729912096556, 2.6146112239752224, 1.7164197479589602]       # TR:204-209 with the default exponent
OPTIMIZER_CFG = "Adam lr=1e-3 eps=1e-8 wd=1e-6 (script defaults TR:566-568)"


def shared_config(world):
    """`config` of the JSON line: identical for our arm and the reference arm (same workload, same optimizer rule)."""
    return dict(workload=WORKLOAD, local_batch=LOCAL_BATCH, global_batch=LOCAL_BATCH * world, optimizer=OPTIMIZER_CFG,
                parallelism="dp%d" % world,
                l2="per-step working set (>7 GB of activations) exceeds the 126 MB L2; no explicit flush")


def load_reference():
    """The UNMODIFIED reference classes (architecture/deeplab_xception.py, utils/losses.py), staged byte for byte under
    baseline/_ref/ by baseline/stage_reference.py, loaded by file path under private names (they never shadow the product's
    `architecture` / `utils` packages).  None when the copy is absent (then the oracle port stands in, kind "port")."""
    import importlib.util
    paths = dict(dx=os.path.join(REF_DIR, "architecture", "deeplab_xception.py"), ls=os.path.join(REF_DIR, "utils", "losses.py"))
    if not all(os.path.exists(v) for v in paths.values()):
        return None
    mods = {}
    for k, v in paths.items():
        spec = importlib.util.spec_from_file_location("_deepcam_reference_" + k, v)
        mods[k] = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mods[k])
    return mods["dx"], mods["ls"]


def synthetic_host_batch(seed, local_batch=None):
    """uniform [0,1) inputs, labels with the class frequencies of the real data (TR:206)."""
    import torch
    n = local_batch or LOCAL_BATCH
    g = torch.Generator().manual_seed(seed)
    x = torch.rand((n, C_IN, H, W), generator=g)
    u = torch.rand((n, H, W), generator=g)
    label = torch.zeros((n, H, W), dtype=torch.long)
    label[u > 0.986267818] = 2
    label[u > 0.986267818 + 0.013274311] = 1
    return x, label


class ReferenceStep:
    """Loop body TR:345-371 on the reference's own classes: forward, fp_loss, zero_grad, backward, Adam step."""

    def __init__(self, device, gpu_library=False):
        import torch
        ref = load_reference()
        self.kind = "reference" if ref is not None else "port"
        self.device = device
        self.gpu_library = gpu_library
        torch.manual_seed(333)
        if ref is not None:
            dxm, lsm = ref
            self.net = dxm.DeepLabv3_plus(n_input=C_IN, n_classes=N_CLASSES, os=16, pretrained=False, _print=False).to(device).train()
            self.loss_fn = lsm.fp_loss
            if gpu_library:
                self.net = self.net.to(memory_format=torch.channels_last)
            kw = dict(fused=True) if gpu_library else {}
            self.opt = torch.optim.Adam(self.net.parameters(), lr=1e-3, eps=1e-8, weight_decay=1e-6, **kw)
        else:
            sys.path.insert(0, os.path.join(REPO, "oracle"))
            import deepcam_oracle as O
            self.O = O
            self.st = O.TrainState({k: v.to(device) for k, v in O.init_state_dict(C_IN, N_CLASSES, 16, seed=333).items()})

    def step(self, x, label):
        import torch
        if self.kind == "port":
            if self.gpu_library:
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return self.st.step(x, label)[0]
            return self.st.step(x, label)[0]
        cw = CLASS_WEIGHTS
        if self.gpu_library:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = self.net.forward(x)
                loss = self.loss_fn(out, label, weight=cw, fpw_1=cw[1], fpw_2=cw[2])
        else:
            out = self.net.forward(x)
            loss = self.loss_fn(out, label, weight=cw, fpw_1=cw[1], fpw_2=cw[2])
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return loss


def run_reference(args):
    """Reference arm: the reference's own training step (its unmodified classes from baseline/_ref; the oracle port only if
    that copy is missing) on the host cores, fp32, FULL-SIZE 768x1152 tiles at local batch 2 - the same config as our arm.
    Rank 0 only; the other ranks exit without work."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    # torchrun exports OMP_NUM_THREADS=1 for its workers; this arm is the only process doing work, so it takes every host core
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    if torch.get_num_threads() < avail:
        torch.set_num_threads(avail)
    cores = torch.get_num_threads()
    ref = ReferenceStep(torch.device("cpu"))
    x, label = synthetic_host_batch(333)
    t0 = time.time()
    ref.step(x, label)                      # cold step = first warm-up step, also the calibration
    t_cal = time.time() - t0
    budget = float(os.environ.get("DEEPCAM_REF_BUDGET_S", "300"))
    warm = max(args.warmup - 1, 0)
    steps = args.steps
    note = ""
    if t_cal * (warm + steps) > budget:      # never crop the tile: run fewer full-size steps instead, and say so
        warm = min(warm, 1)
        steps = max(2, min(steps, int(budget / t_cal) - warm))
        note = "; %d of the requested %d timed steps fit the %.0f s budget" % (steps, args.steps, budget)
    for _ in range(warm):
        ref.step(x, label)
    t0 = time.time()
    for _ in range(steps):
        loss = ref.step(x, label)
    dt = time.time() - t0
    value = LOCAL_BATCH * steps / dt
    sample = "%d full-size steps (N=%d, %dx%dx%d, fp32, Adam) of the %s on %d host threads%s" % (
        steps, LOCAL_BATCH, H, W, C_IN, "unmodified reference classes (baseline/_ref)" if ref.kind == "reference" else
        "oracle port (baseline/_ref absent)", cores, note)
    line = dict(metric=METRIC, value=value, unit="samples/s", n_gpus=args.gpus, steps=steps, warmup=warm + 1,
                ms_per_step=1000.0 * dt / steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference", config=shared_config(max(1, args.gpus)),
                cpu_baseline=dict(value=value, unit="samples/s", cores=cores, kind=ref.kind, sample=sample),
                e2e=dict(value=value, unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                loss=float(loss.detach()) if hasattr(loss, "detach") else float(loss))
    print(json.dumps(line), file=_JSON_OUT, flush=True)


def cpu_baseline_sample():
    """Two full-size reference steps (N=2, 768x1152x16, fp32) on the host cores; the second (warm) one is the value."""
    import torch
    cores = torch.get_num_threads()
    ref = ReferenceStep(torch.device("cpu"))
    x, label = synthetic_host_batch(333)
    t0 = time.time()
    ref.step(x, label)
    t1 = time.time()
    ref.step(x, label)
    dt = time.time() - t1
    return dict(value=LOCAL_BATCH / dt, unit="samples/s", cores=cores, kind=ref.kind,
                sample="1 warm full-size step after 1 cold one (%.1f s / %.1f s), N=%d, %dx%dx%d, fp32, Adam, %d threads"
                       % (dt, t1 - t0, LOCAL_BATCH, H, W, C_IN, cores))


def gpu_library_step_time(dev, steps, warmup):
    """Measurement-only GPU anchor: the reference model on this B200 through stock torch/cuDNN - bf16 autocast,
    channels_last, cudnn.benchmark, fused torch Adam - same batch, same loop body.  Never part of the product path."""
    import torch
    torch.backends.cudnn.benchmark = True
    ref = ReferenceStep(dev, gpu_library=True)
    x, label = synthetic_host_batch(333)
    x = x.to(dev).contiguous(memory_format=torch.channels_last)
    label = label.to(dev)
    for _ in range(max(warmup, 3)):
        ref.step(x, label)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = ref.step(x, label)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    peak_gb = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    out = dict(value=LOCAL_BATCH / (ms / 1000.0), unit="samples/s", ms_per_step=ms, kind=ref.kind, loss=float(loss),
               how="reference DeepLabv3_plus + fp_loss + torch.optim.Adam(fused=True) under torch.autocast(bfloat16), channels_last, "
                   "cudnn.benchmark, torch %s / cuDNN %s; inputs resident in HBM" % (torch.__version__, torch.backends.cudnn.version()),
               peak_mem_gib=peak_gb)
    del ref
    torch.cuda.empty_cache()
    return out


def run_torch_gpu(args):
    """`--impl torch_gpu`: prints the library anchor as its own JSON line (rank 0 / one GPU)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    r = gpu_library_step_time(dev, args.steps, args.warmup)
    line = dict(metric=METRIC, value=r["value"], unit="samples/s", n_gpus=1, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                data="synthetic", impl="torch_gpu", config=shared_config(1), gpu_library_baseline=r)
    print(json.dumps(line), file=_JSON_OUT, flush=True)


# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from architecture import deeplab_xception as dx
    from utils import losses
    from deepcam_b200 import _lib, ops

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(args.warmup, 3)
    precision = os.environ.get("DEEPCAM_B200_PRECISION", "bf16")

    torch.manual_seed(333)
    net = dx.DeepLabv3_plus(n_input=C_IN, n_classes=N_CLASSES, os=16, _print=False)
    net.precision = precision
    net = net.to(dev).train()
    if LOCAL_BATCH == 1:
        # BatchNorm over ONE pooled value per channel cannot train (the reference raises at DX:425-428); the sweep point
        # at local batch 1 evaluates that single layer with its running statistics, as SURVEY 8(d) config 1 prescribes
        net.global_avg_pool[2].eval()
    model = net
    if world > 1:
        from deepcam_b200.parallel import DistributedDataParallel
        model = DistributedDataParallel(net, bucket_cap_mb=args.bucket_mb, broadcast_buffers=not args.no_broadcast_buffers)
    if args.optimizer == "lars":        # BASELINE.json configs[4]; LARS is not part of the reference
        from deepcam_b200.optim import FusedLARS
        opt = FusedLARS(net.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-6)
        opt_name = "deepcam_b200.optim.FusedLARS momentum=0.9 trust=1e-3"
    elif args.optimizer == "lamb":      # the reference's --optimizer LAMB (apex FusedLAMB, TR:217-218)
        from deepcam_b200.optim import FusedLAMB
        opt = FusedLAMB(net.parameters(), lr=1e-3, eps=1e-8, weight_decay=1e-6)
        opt_name = "deepcam_b200.optim.FusedLAMB (apex FusedLAMB semantics)"
    elif os.environ.get("DEEPCAM_B200_TORCH_ADAM", "0") == "1":
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, eps=1e-8, weight_decay=1e-6)
        opt_name = "torch.optim.Adam"
    else:       # same update rule / state as torch.optim.Adam, one launch for all 301 parameters
        from deepcam_b200.optim import FusedAdam
        opt = FusedAdam(net.parameters(), lr=1e-3, eps=1e-8, weight_decay=1e-6)
        opt_name = "deepcam_b200.optim.FusedAdam (torch.optim.Adam semantics)"
    cw = CLASS_WEIGHTS

    x_host, label_host = synthetic_host_batch(333 + rank)
    x_host, label_host = x_host.pin_memory(), label_host.pin_memory()
    x_dev, label_dev = x_host.to(dev), label_host.to(dev)

    def step(x, label, delay=None, module=None):
        if delay:
            delay()
        out = (module or model).forward(x)
        loss = losses.fp_loss(out, label, weight=cw, fpw_1=cw[1], fpw_2=cw[2])
        opt.zero_grad()
        if delay:
            delay()
        loss.backward()
        opt.step()
        return loss

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        sync_all()
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    for _ in range(warmup):
        step(x_dev, label_dev)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count
    ms = timed(lambda: step(x_dev, label_dev), args.steps)
    launches = _lib.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else {}
    value = world * LOCAL_BATCH * args.steps / (ms / 1000.0)

    # ---- end to end through the public API with host buffers: H2D of the batch + D2H of the loss every step ----
    last = {}
    copy_stream = torch.cuda.Stream(device=dev)
    staged = {}

    def stage():
        # host -> device copy of the NEXT step's batch from pinned memory on a copy stream (what a DataLoader with
        # pin_memory + non_blocking transfers does); it overlaps the compute of the current step
        with torch.cuda.stream(copy_stream):
            xb = x_host.to(dev, non_blocking=True)
            lb = label_host.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        staged["next"] = (xb, lb, ev)

    def e2e_step():
        xb, lb, ev = staged["next"]
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ev)
        xb.record_stream(cur)
        lb.record_stream(cur)
        stage()                                       # exactly one batch is copied per step
        loss = step(xb, lb)
        # device -> host read of the step's result, every step: the scalar is copied into pinned memory right behind the
        # step's kernels and the host consumes it one step later (asynchronous logging), so the device never idles while
        # Python enqueues the next step; --e2e-sync-loss reads it with .item() instead (the device then waits for the host)
        if args.e2e_sync_loss:
            last["loss"] = loss.item()
            return
        slot = loss_slots[e2e_state["i"] & 1]
        slot[0].copy_(loss.detach().reshape(1), non_blocking=True)
        slot[1].record()
        e2e_state["i"] += 1
        prev = loss_slots[e2e_state["i"] & 1]
        if e2e_state["i"] > 1:
            prev[1].synchronize()
            last["loss"] = float(prev[0][0])

    def e2e_drain():
        if not args.e2e_sync_loss and e2e_state["i"] > 0:
            cur = loss_slots[(e2e_state["i"] - 1) & 1]
            cur[1].synchronize()
            last["loss"] = float(cur[0][0])

    loss_slots = [(torch.zeros(1, dtype=torch.float32).pin_memory(), torch.cuda.Event()) for _ in range(2)]
    e2e_state = {"i": 0}
    stage()
    for _ in range(2):
        e2e_step()

    def e2e_run():
        e2e_step()

    ms_e2e = timed(e2e_run, args.steps)            # timed() synchronises the device on both sides
    e2e_drain()
    e2e_value = world * LOCAL_BATCH * args.steps / (ms_e2e / 1000.0)
    h2d = x_host.numel() * 4 + label_host.numel() * 8

    # ---- per-kernel-class roofline pass: CUDA events around every launch, same process, right after timing ----
    roofline = None
    classes = None
    if rank == 0 and not args.no_profile:
        peaks = _peaks()
        # per-launch CUDA events need the eager engine (inside a CUDA graph there is nothing to put events around);
        # a spin kernel in front of each phase lets the host run ahead, so the event pairs see back-to-back device
        # execution and not the Python launch overhead
        graphs_env = os.environ.get("DEEPCAM_B200_GRAPHS")
        os.environ["DEEPCAM_B200_GRAPHS"] = "0"
        spin = lambda: torch.cuda._sleep(int(0.045 * 1.9e9))
        # rank 0 only: run the bare module without the gradient exchange (the other ranks are not in this pass, a
        # collective here would never complete); kernels and shapes are those of the timed step
        saved_sync = getattr(net, "_dc_grad_sync", None)
        net._dc_grad_sync = None
        step(x_dev, label_dev, module=net)           # eager warm-up (weight caches of the eager path)
        prof = ops.Profiler(keep_launchers=True)
        ops.set_profiler(prof)
        psteps = 2
        for _ in range(psteps):
            step(x_dev, label_dev, delay=spin, module=net)
            torch.cuda.synchronize()
        ops.set_profiler(None)
        net._dc_grad_sync = saved_sync
        if graphs_env is None:
            del os.environ["DEEPCAM_B200_GRAPHS"]
        else:
            os.environ["DEEPCAM_B200_GRAPHS"] = graphs_env
        classes = prof.summary()
        tensor_kinds = ("conv_gemm_tc", "conv_wgrad_tc", "conv_gemm_simt", "conv_wgrad_simt")
        # Second clock for the same launches: every distinct (kernel, layer shape) of the step replayed 10x back to back inside
        # a CUDA graph - the way it executes in the timed region.  The per-launch event pairs of the eager pass add the
        # event/launch overhead (5-7 us) to every launch, which is half the duration of the 13 us middle-flow kernels.
        gtimes = prof.graph_times(reps=10)
        for cname, cd in classes.items():
            tot, ok = 0.0, True
            for k, v in prof.by_tag({cname}).items():
                if k in gtimes:
                    tot += gtimes[k] * v["launches"] / 1000.0
                else:
                    ok = False
            cd["ms_graph"] = tot if ok else None
        top = max(classes.items(), key=lambda kv: kv[1]["ms"])
        name, d = top
        # the dominant kernel = the (kernel, layer shape) with the largest time inside the dominant class; its DRAM traffic
        # per launch comes from the committed ncu --set full capture of the same shape (profiles/ncu_traffic.json)
        shapes_top = prof.by_tag({name})
        shape_name, shape_d = max(shapes_top.items(), key=lambda kv: kv[1]["ms"])
        traffic = traffic_l2 = None
        try:
            with open(os.path.join(REPO, "profiles", "ncu_traffic.json")) as fh:
                ent = json.load(fh).get(shape_name.replace(" +bnstats", ""))
            if ent:
                traffic = ent["dram_bytes_per_launch"]
                traffic_l2 = ent.get("l2_to_sm_bytes_per_launch")
        except Exception:
            traffic = None
        sec = d["ms"] / 1000.0
        gsec = d["ms_graph"] / 1000.0 if d.get("ms_graph") else None
        if name in tensor_kinds:
            ach = d["flops"] / sec / 1e12
            roofline = dict(kernel=name, bound="tensor", achieved=ach, peak=peaks["tf_sustained"], unit="TFLOP/s",
                            frac=ach / peaks["tf_sustained"], traffic=None)
            if gsec:
                roofline["in_graph"] = dict(achieved=d["flops"] / gsec / 1e12, frac=d["flops"] / gsec / 1e12 / peaks["tf_sustained"],
                                            class_ms_per_step=d["ms_graph"] / psteps)
        else:
            ach = d["bytes"] / sec / 1e9
            roofline = dict(kernel=name, bound="hbm", achieved=ach, peak=peaks["hbm"], unit="GB/s", frac=ach / peaks["hbm"],
                            traffic=None)
            if gsec:
                roofline["in_graph"] = dict(achieved=d["bytes"] / gsec / 1e9, frac=d["bytes"] / gsec / 1e9 / peaks["hbm"],
                                            class_ms_per_step=d["ms_graph"] / psteps)
        if "in_graph" in roofline:
            # headline = the in-graph clock (how the launches execute in the timed region); the per-launch event pairs stay
            # beside it as `eager_events` (they swing with the host's launch rate: 0.30-0.37 between two runs of one build)
            ig = roofline.pop("in_graph")
            roofline["eager_events"] = dict(achieved=roofline["achieved"], frac=roofline["frac"], class_ms_per_step=d["ms"] / psteps,
                                            how="CUDA event pair around every launch of the class in %d instrumented eager steps "
                                                "(includes ~5-7 us of event + launch overhead per launch)" % psteps)
            roofline["achieved"], roofline["frac"] = ig["achieved"], ig["frac"]
            roofline["class_ms_per_step_in_graph"] = ig["class_ms_per_step"]
            roofline["timing"] = ("every distinct (kernel, layer shape) launch of the class replayed 10x back to back inside a CUDA "
                                  "graph, one CUDA event pair per replay, on the launching stream, right after the timed region; "
                                  "class time = sum over shapes of launches x that duration")
        ssec = shape_d["ms"] / 1000.0
        roofline["traffic"] = traffic
        roofline["traffic_l2_to_sm"] = traffic_l2      # l1tex__m_xbar2l1tex_read_bytes.sum of the same capture: what the SMs pulled from L2
        roofline["dominant_shape"] = dict(
            kernel=shape_name, launches_per_step=shape_d["launches"] / psteps, us_per_launch=1000.0 * shape_d["ms"] / shape_d["launches"],
            algorithmic_bytes_per_launch=shape_d["bytes"] / shape_d["launches"],
            algorithmic_flops_per_launch=shape_d["flops"] / shape_d["launches"],
            achieved_tflops=shape_d["flops"] / ssec / 1e12 if ssec > 0 else None,
            achieved_gbs=shape_d["bytes"] / ssec / 1e9 if ssec > 0 else None,
            traffic_note="`traffic` = dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this shape "
                         "(cold caches; written lines still in L2 at kernel end are not counted by the counter, so it can read below "
                         "the algorithmic bytes); `traffic_l2_to_sm` = L2->SM bytes of the same capture, the figure that shows operand "
                         "re-fetch (4.1x the algorithmic bytes for the 728x728 layer: every CTA pulls its own copy of the weights)")
        roofline.update(peak_source=peaks["source"] + (" sustained" if name in tensor_kinds else " copy"),
                        launches_per_step=d["launches"] / psteps, avg_launch_us=1000.0 * d["ms"] / d["launches"],
                        class_ms_per_step=d["ms"] / psteps,
                        how="CUDA events around every launch of the class in %d instrumented eager steps after the timed region "
                            "(the timed region itself replays CUDA graphs)" % psteps)
        out_dir = os.path.join(REPO, "gpurun_out")
        try:
            os.makedirs(out_dir, exist_ok=True)
            table = {}
            for k, v in sorted(classes.items(), key=lambda kv: -kv[1]["ms"]):
                s = v["ms"] / 1000.0
                table[k] = dict(launches_per_step=v["launches"] / psteps, ms_per_step=v["ms"] / psteps,
                                tflops=v["flops"] / s / 1e12 if s > 0 else 0.0, gbs=v["bytes"] / s / 1e9 if s > 0 else 0.0,
                                ms_per_step_in_graph=(v["ms_graph"] / psteps) if v.get("ms_graph") else None,
                                flops_per_step=v["flops"] / psteps, bytes_per_step=v["bytes"] / psteps)
            shapes = {}
            for k, v in sorted(prof.by_tag(set(classes)).items(), key=lambda kv: -kv[1]["ms"]):
                s = v["ms"] / 1000.0
                shapes[k] = dict(launches_per_step=v["launches"] / psteps, ms_per_step=v["ms"] / psteps,
                                 us_per_launch=1000.0 * v["ms"] / v["launches"], tflops=v["flops"] / s / 1e12 if s > 0 else 0.0,
                                 gbs=v["bytes"] / s / 1e9 if s > 0 else 0.0, us_per_launch_in_graph=gtimes.get(k),
                                 flops_per_launch=v["flops"] / v["launches"], bytes_per_launch=v["bytes"] / v["launches"])
            with open(os.path.join(out_dir, "bench_kernel_classes.json"), "w") as fh:
                json.dump(dict(ms_per_step_timed=ms / args.steps, classes=table, conv_shapes=shapes, peaks=peaks), fh, indent=1)
        except Exception:
            pass

    # ---- N > 1: exposed communication = step time with the gradient exchange minus the same ranks stepping without it ----
    comm = None
    if world > 1:
        saved_sync, saved_key = net._dc_grad_sync, getattr(net, "_dc_plan_key", None)
        net._dc_grad_sync = None
        net._dc_plan_key = "no_grad_exchange"           # a plan of its own: one backward graph, no bucket segments
        for _ in range(3):                               # eager, capture, replay
            step(x_dev, label_dev, module=net)
        ms_nc = timed(lambda: step(x_dev, label_dev, module=net), args.steps)
        net._dc_grad_sync, net._dc_plan_key = saved_sync, saved_key
        comm = dict(ms_per_step_without_exchange=ms_nc / args.steps, exposed_comm_ms=(ms - ms_nc) / args.steps,
                    how="same ranks, same batches, bare module (no bucketed all-reduce, no buffer broadcast), max over ranks; "
                        "the difference to ms_per_step is everything the exchange costs: exposed NCCL time, SM contention, "
                        "per-bucket graph segments, the buffer broadcast")

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_baseline_sample()
    gpu_lib = None
    if rank == 0 and world == 1 and not args.no_gpu_baseline:
        try:
            gpu_lib = gpu_library_step_time(dev, min(args.steps, 10), 3)
        except Exception as e:                          # the anchor must never take the bench line down
            gpu_lib = dict(unavailable=repr(e)[:200])

    if rank == 0:
        peaks = _peaks()
        line = dict(metric=METRIC, value=value, unit="samples/s", n_gpus=world, steps=args.steps, warmup=warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="bf16" if precision == "bf16" else "f32", data="synthetic",
                    config=dict(shared_config(world), **({} if args.optimizer == "adam" else {"optimizer": opt_name})),
                    optimizer_impl=opt_name,
                    e2e=dict(value=e2e_value, unit="samples/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=4,
                             ms_per_step=ms_e2e / args.steps, loss=last.get("loss"),
                             loss_readback=".item() every step" if args.e2e_sync_loss else
                             "4-byte copy into pinned memory every step, consumed by the host one step later"),
                    gpu_launches=launches, clocks=clocks,
                    tensor_frac_of_step=(FWD_BWD_GFLOP_PER_SAMPLE * value / world / 1000.0) / peaks["tf_sustained"],
                    roofline=roofline, cpu_baseline=cpu_base, gpu_library_baseline=gpu_lib, comm=comm)
        from deepcam_b200 import engine as _engine
        if _engine.deterministic():
            line["deterministic"] = True      # DEEPCAM_B200_DETERMINISTIC=1: two-stage weight-gradient reductions (bit-reproducible)
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.barrier()              # rank 0 may still be in its instrumented pass: leave together
        dist.destroy_process_group()


def run_eval(args):
    """BASELINE.json configs[3]: the validation loop TR:423-512 - eval() + no_grad forward, fp_loss, argmax (TR:458) and
    compute_score (TR:459) per batch, per-rank sums, three scalar all-reduces at the end (TR:490-492) - over a synthetic
    validation set sharded by rank.  The IoU of every batch comes out of ONE fused kernel pass over the logits
    (utils.argmax_score: argmax + integer tp/fp/fn + the three divisions), bit-exact with UT:32-60."""
    import torch
    import torch.distributed as dist
    from architecture import deeplab_xception as dx
    from utils import losses, utils as dcutils
    from deepcam_b200 import _lib
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(333)
    net = dx.DeepLabv3_plus(n_input=C_IN, n_classes=N_CLASSES, os=16, _print=False).to(dev)
    # random-init weights with the initial running statistics (0, 1) blow the eval-mode activations up; ten train-mode forward
    # passes on synthetic batches settle the running statistics first (not timed; what a few training steps would have done)
    xc, _ = synthetic_host_batch(4000 + rank, 2)
    xc = xc.to(dev)
    net.train()
    with torch.no_grad():
        for _ in range(10):
            net.forward(xc)
    del xc
    net.eval()
    b = args.eval_batch
    nbuf = 4                                         # distinct batches cycled through (each 113 MB: beyond L2 together with the activations)
    batches = []
    for i in range(nbuf):
        x, label = synthetic_host_batch(5000 + rank * nbuf + i, b)
        batches.append((x.to(dev), label.to(dev)))
    iters = max(1, args.eval_samples // b)
    cw = CLASS_WEIGHTS
    sums = torch.zeros(3, device=dev)                # count_sum_val, loss_sum_val, iou_sum_val (TR:431-433)

    def one(i):
        x, label = batches[i % nbuf]
        with torch.no_grad():
            out = net.forward(x)
            loss = losses.fp_loss(out, label, weight=cw, fpw_1=cw[1], fpw_2=cw[2])
            iou = dcutils.argmax_score(out, label, N_CLASSES)
            sums[0] += 1.0
            sums[1] += loss
            sums[2] += iou

    for i in range(max(args.warmup, 3)):
        one(i)
    sums.zero_()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        one(i)
    if world > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)  # the three reductions of TR:490-492 as one
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count - l0
    clocks = sampler.stop() if rank == 0 else {}
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    if rank == 0:
        total = world * iters * b
        host = sums.cpu()
        line = dict(metric="eval_samples_per_s_768x1152x16", value=total / (ms / 1000.0), unit="samples/s", n_gpus=world,
                    steps=iters, warmup=max(args.warmup, 3), ms_per_step=ms / iters, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype=os.environ.get("DEEPCAM_B200_PRECISION", "bf16"), data="synthetic",
                    config=dict(workload="configs[3]: eval path, forward + fp_loss + argmax + IoU over a synthetic validation set "
                                         "(%d samples per rank, batch %d), validation loop TR:423-512" % (iters * b, b),
                                batch=b, samples=total, parallelism="dp%d (validation set sharded by rank)" % world,
                                l2="%d distinct batches cycled; activations exceed L2" % nbuf),
                    eval_loss=float(host[1] / host[0]), eval_iou=float(host[2] / host[0]), gpu_launches=launches, clocks=clocks)
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_gpu"])
    ap.add_argument("--mode", default="train", choices=["train", "eval"],
                    help="train = the headline metric; eval = BASELINE.json configs[3] (validation loop TR:423-512)")
    ap.add_argument("--eval-samples", type=int, default=64, help="--mode eval: validation samples per rank")
    ap.add_argument("--eval-batch", type=int, default=1, help="--mode eval: batch size (the reference validates at 1, TR:302-306)")
    ap.add_argument("--bucket-mb", type=float, default=64.0, help="N > 1: gradient bucket size of the data-parallel wrapper")
    ap.add_argument("--no-broadcast-buffers", action="store_true", help="N > 1: skip the per-forward BatchNorm buffer broadcast (C3)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--e2e-sync-loss", action="store_true", help="e2e leg: read the loss with .item() every step")
    ap.add_argument("--optimizer", default="adam", choices=["adam", "lamb", "lars"],
                    help="adam = the headline configuration (script default, TR:566-568); lamb / lars: configs[4] sweep")
    ap.add_argument("--local-batch", type=int, default=LOCAL_BATCH,
                    help="samples per GPU (BASELINE.json configs[4] sweep; the headline metric is quoted at the default, 2)")
    args = ap.parse_args()
    if args.local_batch != LOCAL_BATCH:
        globals()["LOCAL_BATCH"] = args.local_batch
        globals()["WORKLOAD"] = WORKLOAD.replace("local batch 2", "local batch %d (configs[4] sweep point)" % args.local_batch)
    # stdout carries exactly ONE JSON line: libraries that print to file descriptor 1 (NCCL's version banner) are sent to
    # stderr for the duration of the run, the JSON line goes to the saved descriptor
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch_gpu":
        run_torch_gpu(args)
    elif args.mode == "eval":
        run_eval(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
