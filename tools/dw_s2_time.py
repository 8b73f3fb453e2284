import sys, torch
sys.path.insert(0, "mlperf-deepcam_b200")
from deepcam_b200.backend import CudaBackend, DwSpec
dev = torch.device("cuda:0")
be = CudaBackend(torch.bfloat16, dev)
for (h, w_, c) in [(384, 576, 128), (192, 288, 256), (96, 144, 728)]:
    x = torch.randn(2, h, w_, c, device=dev).bfloat16(); out = torch.empty(2, h // 2, w_ // 2, c, device=dev, dtype=torch.bfloat16); dy = torch.randn_like(out); dx = torch.empty_like(x)
    w = torch.nn.Parameter(torch.randn(c, 1, 3, 3, device=dev)); spec = DwSpec("dw", w, 2, 1); wg = torch.zeros(c, 1, 3, 3, device=dev)
    for name, fn in [("fwd", lambda: be.dw_fwd(x, spec, out)), ("bwd_data", lambda: be.dw_bwd_data(dy, spec, dx, False)), ("bwd_weight", lambda: be.dw_bwd_weight(x, dy, spec, wg))]:
        for _ in range(3): fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20): fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        print("dw s2 %dx%dx%d %s %.1f us" % (h, w_, c, name, e0.elapsed_time(e1) * 50))
