"""Compare block-output gradients of the CUDA fp32 path with the fp64 oracle (debug aid)."""
import os, sys
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "semi-seg-ecg_b200", "src")); sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np, torch
from helpers import O, TRAIN_CFG, batches, model_cfg, rel_err
from algorithms.base import init_model_from_cfg
from semiseg_b200 import _lib
from semiseg_b200.trainer import get_engine

mode = sys.argv[1] if len(sys.argv) > 1 else "sup"
torch.manual_seed(0)
cfgm = model_cfg(1, 64, 64, 128, 0.0)
model = init_model_from_cfg(cfgm)
init = {k: v.detach().clone() for k, v in model.state_dict().items()}
model = model.to("cuda")
cfg = dict(TRAIN_CFG, conf_thresh=0.3)
(lab, unl), = batches(400, 1, 2, 2, 1, 2500)
tr = O.OracleTrainer(init, O.Arch(num_leads=1, dropout_ratio=0.0), cfg, dtype=torch.float64)
if mode == "sup":
    x, y = torch.cat((lab["ecg"], unl["ecg_aug"])), torch.cat((lab["target"], lab["target"]))
    tr.supervised_step(x, y, 3e-4, want_taps=True)
    eng = get_engine("supervised", model, None, 4, 0, 2500, _lib.F32, cfg, use_graph=False, algo=_lib.ALGO_SIMT)
    eng.load_batch(x, y)
else:
    tr.fixmatch_step(lab["ecg"], lab["target"], unl["ecg"], unl["ecg_aug"], 3e-4, want_taps=True)
    eng = get_engine("fixmatch", model, None, 2, 2, 2500, _lib.F32, cfg, use_graph=False, algo=_lib.ALGO_SIMT)
    eng.load_batch(lab["ecg"], lab["target"], unl["ecg"], unl["ecg_aug"])
eng.plan_s.debug = {}
eng.step(3e-4)
print(eng.read_stats())
# forward check of the block output at the suspicious spot
plan = eng.plan_s
bi = [i for i, b in enumerate(plan.lay.blocks) if b.prefix == "backbone.layer3.0"][0]
mine_out = plan.to_ncl(plan.blk_bufs[bi]["out"], plan.g_stage[2]).cpu().double()
ref_out = tr.taps["backbone.layer3.0"].detach()
dd = (mine_out - ref_out).abs()
print("layer3.0 out: max abs diff", float(dd.max()), "rel", rel_err(mine_out, ref_out))
mism = ((mine_out > 0) != (ref_out > 0))
print("relu-mask mismatches:", int(mism.sum()), "at", mism.nonzero()[:10].tolist())
for (b, c, t) in mism.nonzero()[:8].tolist():
    print("   ", b, c, t, "mine", float(mine_out[b, c, t]), "ref", float(ref_out[b, c, t]))
for name, G in eng.plan_s.debug.items():
    ref = tr.taps[name].grad
    d = (G.cpu().double() - ref)
    if rel_err(G, ref) < 1e-4: continue
    per_pos = d.pow(2).sum(dim=(0, 1)).sqrt() / (ref.pow(2).sum(dim=(0, 1)).sqrt() + 1e-30)
    worst = torch.topk(per_pos, 3)
    per_b = d.pow(2).sum(dim=(1, 2)).sqrt() / ref.pow(2).sum(dim=(1, 2)).sqrt()
    if "layer3.0" in name:
        b = int(torch.argmax(per_b)); pos = int(worst.indices[0])
        ch = d[b, :, pos].abs()
        print("   bad column: nonzero-diff channels", int((ch > 1e-9).sum()), "of", ch.numel(), "max diff", float(ch.max()), "ref max", float(ref[b, :, pos].abs().max()),
              "mine[:4]", G[b, :4, pos].tolist(), "ref[:4]", ref[b, :4, pos].tolist())
    print(f"{name:24s} err {rel_err(G, ref):.2e}  worst positions {worst.indices.tolist()} {[f'{v:.1e}' for v in worst.values.tolist()]} per-sample {[f'{v:.1e}' for v in per_b.tolist()]}")
