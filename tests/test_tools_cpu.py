"""Host-side tooling that needs no GPU: the ncu-CSV -> reference-notebook export and the data-parallel
bucket layout (tail bucket)."""
import csv
import os
import re
import subprocess
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_analysis_export_matches_reference_notebook_conventions(tmp_path):
    src = os.path.join(REPO, "profiles", "r02_launches_step_graph.csv")
    out = str(tmp_path)
    subprocess.run([sys.executable, os.path.join(REPO, "tools", "analysis_export.py"), src, "--batchsize", "2", "--pass", "training",
                    "--out", out], check=True)
    fn = os.path.join(out, "deepcam.batchsize_2.pass_training.metrics.csv")
    # the reference takes batch size and pass from the file name (analysis/utils.py:31-42)
    assert int(re.match(r'.*\.batchsize_(.*?)\.', fn).groups()[0]) == 2
    assert re.match(r'.*\.pass_(.*?)\.', fn).groups()[0] == "training"
    rows = list(csv.DictReader(open(fn)))
    assert rows and set(rows[0]) == {"Name", "Metric Name", "Invocations", "Metric Value"}        # import_nsight_metric's frame
    ov = list(csv.DictReader(open(fn.replace(".metrics.", ".overview."))))
    assert set(ov[0]) == {"Name", "Time", "Invocations", "Time Avg"}                              # import_nsight_overview's frame
    names = {r["Name"] for r in ov}
    assert any("conv_gemm_tc2_kernel" in n for n in names) and any("conv_wgrad_tc_kernel" in n for n in names)
    total_ms = sum(float(r["Time"]) for r in ov)
    assert 5.0 < total_ms < 200.0
    assert abs(float(ov[0]["Time"]) / int(ov[0]["Invocations"]) - float(ov[0]["Time Avg"])) < 1e-9
    assert os.path.exists(os.path.join(out, "b200_ceilings.json"))


def test_gradient_bucket_layout_covers_the_flat_buffer_with_a_small_tail():
    sys.path.insert(0, os.path.join(REPO, "mlperf-deepcam_b200"))
    from architecture import deeplab_xception as dx
    from deepcam_b200.engine import GradStore
    from deepcam_b200.parallel import _GradSync

    class Owner:
        bucket_cap_elems = int(64 * 2 ** 20 // 4)
        tail_cap_elems = int(8 * 2 ** 20 // 4)

    net = dx.DeepLabv3_plus(16, 3, 16, _print=False)
    gs = GradStore(list(net.parameters()), torch.device("cpu"))
    sync = _GradSync(Owner())
    sync._total = gs.total
    sync._layout(gs)
    spans = sorted((a, b) for a, b, _ in sync.buckets)
    assert spans[0][0] == 0 and spans[-1][1] == gs.total
    assert all(spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1))                      # contiguous, no overlap
    assert sorted(i for _, _, idx in sync.buckets for i in idx) == list(range(len(gs.params)))    # every parameter once
    # launch order = reverse parameter order; the last bucket (earliest layers, nothing left to overlap with) is the small one
    last = sync.buckets[-1]
    assert last[0] == 0 and (last[1] - last[0]) * 4 <= 8 * 2 ** 20
    assert all((b - a) * 4 <= 64 * 2 ** 20 + 4 * 2 ** 20 for a, b, _ in sync.buckets)
