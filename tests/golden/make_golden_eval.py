"""Golden vectors for the evaluation path, made by running the UNMODIFIED reference
`algorithms.base.evaluate` (base.py:184-245) on CPU (TEST INFRASTRUCTURE; build container only):

    python tests/golden/make_golden_eval.py

torchmetrics (1.5.2 in requirements.txt) is not installable here, so the `metric_fn` ARGUMENT handed to the reference's
evaluate is oracle/eval_oracle.MetricCollection([MeanIoU]) -- the restated torchmetrics algorithm; everything else
(eval forward, loss, soft-max, arg-max, one-hot encoding, meters) is the reference's own code.  Case H: the tiny
network after a few train-mode forwards (non-trivial running statistics), three batches of different sizes (5, 5, 2)
so that the sample-weighted loss mean and the batch-mean-of-batch-means IoU differ from their naive versions.
Stores the state dict, the data seeds and the returned stats / metrics / outputs / labels in eval_vectors.npz."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)
from oracle.ref_harness import ListLoader, import_reference  # noqa: E402
import make_golden as mg  # noqa: E402
from oracle import eval_oracle  # noqa: E402

SIZES = (5, 5, 2)


def eval_batches():
    out = []
    for i, n in enumerate(SIZES):
        lab, _ = mg.synthetic.make_batch(700 + i, n, 1, mg.TINY["num_leads"], mg.TINY["L"])
        out.append({k: torch.from_numpy(v) for k, v in lab.items()})
    return ListLoader(out)


def main():
    R = import_reference()
    torch.set_num_threads(4)
    T = mg.TINY
    cfg = mg.model_cfg(T["num_leads"], T["stem_channels"], T["base_channels"], T["head_channels"], 0.0)
    torch.manual_seed(21)
    model = R.base.init_model_from_cfg(cfg)
    model.train()
    with torch.no_grad():
        for i in range(3):
            model(torch.from_numpy(mg.synthetic.make_batch(650 + i, 4, 1, T["num_leads"], T["L"])[0]["ecg"]))
        # make the classes compete: a random bias on the classifier so that the arg-max is not one class everywhere
        model.decode_head.cls_seg.weight.mul_(8.0)
    out = {}
    mg.put(out, "H/model", mg.to_np(model.state_dict()))
    for tag, kw in (("H", {}), ("Hnb", {"include_background": False}), ("Hpc", {"per_class": True})):
        metric_fn = eval_oracle.MetricCollection([eval_oracle.MeanIoU(4, **kw)])
        stats, metrics, outputs, labels = R.base.evaluate(model, eval_batches(), torch.device("cpu"), metric_fn, use_amp=False)
        mg.put(out, f"{tag}/stats", {k: np.float64(v) for k, v in stats.items()})
        mg.put(out, f"{tag}/metrics", {k: np.float64(v) for k, v in metrics.items()})
        if tag == "H":
            out["H/outputs"] = outputs.numpy().astype(np.float32)
            out["H/labels_onehot"] = labels.numpy().astype(np.uint8)
    out["H/sizes"] = np.array(SIZES)
    out["H/data_seed"] = np.int64(700)
    path = os.path.join(HERE, "eval_vectors.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1e6:.2f} MB")
    print({k: float(out[k]) for k in out if "/stats/" in k or "/metrics/" in k})
    pred = outputs.argmax(1)
    print("predicted class histogram", np.bincount(pred.numpy().ravel(), minlength=4))


if __name__ == "__main__":
    main()
