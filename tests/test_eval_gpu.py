"""Evaluation path on the GPU: ssb_eval_metrics against a torch restatement, and
algorithms.base.evaluate against the reference's evaluate (golden case H)."""
import numpy as np
import pytest
import torch

from helpers import TINY_ARCH, eval_batches, group, model_cfg, rel_err, sd_from  # noqa: F401

pytestmark = pytest.mark.gpu

from semiseg_b200 import _lib  # noqa: E402
from semiseg_b200._lib import call  # noqa: E402
from semiseg_b200.evaluate import mean_iou_from_counts  # noqa: E402

DEV = "cuda"


def st():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("B,Lin,L,ncls,align", [(3, 79, 2500, 4, 0), (2, 10, 37, 3, 0), (1, 157, 5000, 4, 1), (5, 8, 8, 8, 0),
                                                (2, 1, 9, 2, 0)])
def test_eval_metrics_kernel(B, Lin, L, ncls, align):
    torch.manual_seed(B * 1000 + L)
    low = (3 * torch.randn(B, Lin, ncls, device=DEV)).contiguous()
    y = torch.randint(0, ncls, (B, L), device=DEV)
    y[0, : min(5, L)] = -100                        # ignored positions (nn.CrossEntropyLoss ignore_index)
    sums = torch.zeros(2, dtype=torch.float64, device=DEV)
    counts = torch.zeros(B, ncls, 3, dtype=torch.int32, device=DEV)
    probs = torch.full((B, ncls, L), -1.0, device=DEV)
    pred = torch.full((B, L), -1, dtype=torch.int64, device=DEV)
    call("ssb_eval_metrics", low.data_ptr(), y.data_ptr(), sums.data_ptr(), counts.data_ptr(), probs.data_ptr(), pred.data_ptr(),
         B, Lin, L, ncls, align, st())
    # the same call without the optional outputs
    sums2, counts2 = torch.zeros_like(sums), torch.zeros_like(counts)
    call("ssb_eval_metrics", low.data_ptr(), y.data_ptr(), sums2.data_ptr(), counts2.data_ptr(), None, None, B, Lin, L, ncls,
         align, st())
    torch.cuda.synchronize()
    logits = torch.nn.functional.interpolate(low.permute(0, 2, 1).double(), size=L, mode="linear", align_corners=bool(align))
    ref_p = logits.softmax(1)
    assert rel_err(probs, ref_p) < 1e-5
    top2 = ref_p.topk(2, dim=1).values
    decided = (top2[:, 0] - top2[:, 1]) > 1e-5
    assert torch.equal(pred[decided], ref_p.argmax(1)[decided])
    assert torch.equal(pred, probs.argmax(1))            # self-consistent: first maximum of the written probabilities
    ce = torch.nn.functional.cross_entropy(logits, y, ignore_index=-100, reduction="sum")
    assert abs(float(sums[0]) - float(ce)) < 1e-5 * max(1.0, float(ce))
    assert float(sums[1]) == float((y >= 0).sum())
    valid = y >= 0
    for c in range(ncls):
        assert torch.equal(counts[:, c, 0].long(), ((pred == c) & (y == c)).sum(1))
        assert torch.equal(counts[:, c, 1].long(), (pred == c).sum(1))
        assert torch.equal(counts[:, c, 2].long(), ((y == c) & valid).sum(1))
    assert torch.equal(counts, counts2) and torch.allclose(sums, sums2, rtol=1e-12)


def test_eval_metrics_rejects_bad_arguments():
    z = torch.zeros(8, device=DEV)
    lib = _lib.load()
    assert lib.ssb_eval_metrics(z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), None, None, 1, 1, 1, 9, 0, st()) != 0
    assert lib.ssb_eval_metrics(None, z.data_ptr(), z.data_ptr(), z.data_ptr(), None, None, 1, 1, 1, 4, 0, st()) != 0


def _model(g):
    from algorithms.base import init_model_from_cfg
    m = init_model_from_cfg(model_cfg(2, 8, 8, 16, 0.0))
    m.load_state_dict(sd_from(g, "H/model"))
    return m.to(DEV)


@pytest.mark.parametrize("tag,opts", [("H", {}), ("Hnb", {"include_background": False}), ("Hpc", {"per_class": True})])
def test_evaluate_matches_reference(golden_eval, tag, opts):
    """algorithms.base.evaluate (fp32) vs the reference's evaluate: sample-weighted loss, soft-max outputs, one-hot labels
    and MeanIoU in the three metric configurations (batches of sizes 5, 5, 2: two graphs)."""
    from algorithms.base import evaluate
    g = golden_eval
    model = _model(g)
    stats, metrics, outputs, labels = evaluate(model, eval_batches(g), torch.device(DEV), dict(opts), use_amp=False,
                                               return_outputs=True)
    assert abs(stats["loss"] - float(g[f"{tag}/stats/loss"])) < 1e-5 * max(1.0, float(g[f"{tag}/stats/loss"]))
    ref_out = torch.from_numpy(g["H/outputs"])
    assert rel_err(outputs, ref_out) < 1e-5
    assert np.array_equal(labels.numpy().astype(np.uint8), g["H/labels_onehot"])
    top2 = ref_out.topk(2, dim=1).values
    undecided = int(((top2[:, 0] - top2[:, 1]) <= 1e-5).sum())
    same_pred = torch.equal(outputs.argmax(1), ref_out.argmax(1))
    assert same_pred or undecided > 0
    ref = group(g, f"{tag}/metrics")
    assert set(metrics) == set(ref)
    for k in ref:
        tol = 1e-9 if same_pred else 1e-3
        assert abs(metrics[k] - float(ref[k])) < tol, (k, metrics[k], float(ref[k]))
    # without the outputs: same numbers, nothing returned
    s2, m2, o2, l2 = evaluate(model, eval_batches(g), torch.device(DEV), dict(opts), use_amp=False)
    assert o2 is None and l2 is None and s2 == stats and m2 == metrics


def test_evaluate_bf16_and_after_training(golden_eval):
    """bf16 path within the north_star tolerance of the fp32 reference numbers; and evaluate() sees weights changed by
    a training step (the storage-dtype weight copy is refreshed per evaluation)."""
    from algorithms.base import evaluate, train_one_epoch
    from helpers import TRAIN_CFG, batches
    from utils.optimizer import get_optimizer_from_config
    g = golden_eval
    model = _model(g)
    stats, metrics, _, _ = evaluate(model, eval_batches(g), torch.device(DEV), None, use_amp=True)
    ref = float(g["H/stats/loss"])
    assert abs(stats["loss"] - ref) < 2e-2 * ref
    assert abs(metrics["MeanIoU"] - float(g["H/metrics/MeanIoU"])) < 2e-2
    opt = get_optimizer_from_config(TRAIN_CFG, model.parameters())
    data = batches(900, 2, 3, 3, 2, 300)
    train_one_epoch(model, [d[0] for d in data], opt, torch.device(DEV), 30, None, None, True, dict(TRAIN_CFG))
    stats2, _, _, _ = evaluate(model, eval_batches(g), torch.device(DEV), None, use_amp=True)
    assert stats2["loss"] != stats["loss"] and np.isfinite(stats2["loss"])


def test_predict_loader_matches_reference_outputs(golden_eval):
    """the inference loop (reference inference.py:108-119): soft-max outputs of every batch, on the CPU"""
    from semiseg_b200.evaluate import predict_loader
    g = golden_eval
    out = predict_loader(_model(g), [{"ecg": b["ecg"]} for b in eval_batches(g)], torch.device(DEV), use_amp=False)
    assert out.device.type == "cpu" and tuple(out.shape) == g["H/outputs"].shape
    assert rel_err(out, g["H/outputs"]) < 1e-5
    with pytest.raises(RuntimeError):
        predict_loader(_model(g), [], torch.device("cpu"))


def test_mean_iou_from_counts_semantics():
    """a class absent from prediction and target scores 0 (torchmetrics _safe_divide), not NaN and not skipped"""
    counts = torch.tensor([[[5, 10, 5], [0, 0, 0], [0, 3, 0]]], device=DEV)
    s = mean_iou_from_counts(counts)
    assert abs(float(s) - (0.5 + 0.0 + 0.0) / 3) < 1e-12
    s = mean_iou_from_counts(counts, include_background=False, per_class=True)
    assert s.shape == (1, 2) and float(s.sum()) == 0.0
