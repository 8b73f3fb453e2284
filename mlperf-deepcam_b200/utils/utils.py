"""B200-native drop-in for the reference's `utils.utils` (UT = reference src/deepCam/utils/utils.py).

`compute_score(prediction, gt, num_classes, device_id, type, weights)` keeps the reference signature (UT:32).
The reference forms nine masked sums and synchronises the host three times (`union.item()`, UT:55); here one
integer kernel produces the int64 tp/fp/fn counters (bit-exact) and a second tiny kernel evaluates
sum_j iou_j / num_classes in fp32 with the reference's rules (empty union -> 1.0), without a host sync.
`device_id`, `type` and `weights` are unused, exactly as in the reference.
"""
import torch

from deepcam_b200 import ops


def iou_counts(prediction, gt, num_classes, counts=None):
    """int64 tensor [3*num_classes] = tp | fp | fn (UT:43-50).  `counts` may be passed to accumulate."""
    if not prediction.is_cuda:
        raise RuntimeError("deepcam_b200.compute_score needs CUDA tensors (got %s); there is no CPU fallback"
                           % prediction.device)
    gt = gt.type(torch.long)                        # UT:41
    if prediction.dtype != torch.int64:
        prediction = prediction.long()
    if prediction.shape != gt.shape:
        prediction, gt = torch.broadcast_tensors(prediction, gt)
    prediction = prediction.contiguous()
    gt = gt.contiguous()
    if counts is None:
        counts = torch.empty(3 * num_classes, dtype=torch.int64, device=prediction.device)
        ops.fill_zero(counts)
    ops.iou_counts(prediction, gt, num_classes, counts)
    return counts


def compute_score(prediction, gt, num_classes, device_id=None, type="iou", weights=None):
    counts = iou_counts(prediction, gt, num_classes)
    score = torch.empty(1, dtype=torch.float32, device=prediction.device)
    ops.iou_finalize(counts, num_classes, score)
    return score.reshape(())


def argmax_score(logits, gt, num_classes=None, return_predictions=False):
    """Fused eval-path helper: torch.max(logits, 1)[1] (first-max tie rule, TR:458) + compute_score (TR:459)
    in one pass over the logits."""
    n, c, h, w = logits.shape
    num_classes = num_classes or c
    lg = logits.detach()
    if lg.dtype != torch.float32:
        lg = lg.float()
    gt = gt.type(torch.long).contiguous()
    counts = torch.empty(3 * num_classes, dtype=torch.int64, device=lg.device)
    ops.fill_zero(counts)
    pred = torch.empty((n, h, w), dtype=torch.int64, device=lg.device) if return_predictions else None
    ops.argmax_iou(lg.permute(0, 2, 3, 1), gt, num_classes, pred, counts)
    score = torch.empty(1, dtype=torch.float32, device=lg.device)
    ops.iou_finalize(counts, num_classes, score)
    return (score.reshape(()), pred) if return_predictions else score.reshape(())
