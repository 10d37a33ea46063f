#!/usr/bin/env python
"""The "library Blackwell" bar (VERDICT round 1 missing 1): the SAME FixMatch step written the way the
reference writes it -- torch.nn modules, autograd, torch.optim.AdamW, eager mode, cuDNN 9 convolutions -- timed on the
same B200 with CUDA events.  This is what a user gets by running the reference's algorithm on this GPU with stock
PyTorch; the hand-written path has to beat it, not just the CPU.

Not product code and not the oracle: a plain restatement of the published architecture (1-D ResNet-18 encoder with
BasicBlocks, FCN head, linear upsampling; reference call sites: resnet.py:55-72,206-257, fcn_head.py:89-97,
encoder_decoder.py:78-111) and of the step (fixmatch.py:79-138: eval-mode fp32 pseudo-label pass without autocast,
train-mode student pass under autocast, CE + masked CE, GradScaler, AdamW; 3 `.item()` reads and a synchronize per step
like the reference's logging).  Used by bench.py's `library_gpu_baseline` block.
"""
import time

import torch
import torch.nn as nn
import torch.nn.functional as F


class Block(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv1d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm1d(cout)
        self.conv2 = nn.Conv1d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm1d(cout)
        self.down = None
        if stride != 1 or cin != cout:
            self.down = nn.Sequential(nn.Conv1d(cin, cout, 1, stride, bias=False), nn.BatchNorm1d(cout))

    def forward(self, x):
        idt = x if self.down is None else self.down(x)
        o = F.relu(self.bn1(self.conv1(x)), inplace=True)
        o = self.bn2(self.conv2(o))
        return F.relu(o + idt, inplace=True)


class SegNet(nn.Module):
    def __init__(self, leads=1, base=64, stem=64, head=128, ncls=4, p=0.1):
        super().__init__()
        self.stem = nn.Sequential(nn.Conv1d(leads, stem, 7, 2, 3, bias=False), nn.BatchNorm1d(stem), nn.ReLU(inplace=True),
                                  nn.MaxPool1d(3, 2, 1))
        layers, cin = [], stem
        for i in range(4):
            cout = base * 2 ** i
            layers += [Block(cin, cout, 1 if i == 0 else 2), Block(cout, cout, 1)]
            cin = cout
        self.layers = nn.Sequential(*layers)
        self.head = nn.Sequential(nn.Conv1d(cin, head, 3, 1, 1, bias=False), nn.BatchNorm1d(head), nn.ReLU(inplace=True))
        self.drop = nn.Dropout(p)
        self.cls = nn.Conv1d(head, ncls, 1)

    def forward(self, x):
        L = x.shape[2]
        h = self.cls(self.drop(self.head(self.layers(self.stem(x)))))
        return F.interpolate(h, size=L, mode="linear", align_corners=False)


def fixmatch_step(model, opt, scaler, amp_dtype, x, y, uw, us, thr):
    with torch.no_grad():
        model.eval()
        pw = model(uw)
        conf = pw.softmax(dim=1).max(dim=1)[0]
        lab = pw.argmax(dim=1)
    model.train()
    nl = x.shape[0]
    with torch.autocast("cuda", dtype=amp_dtype, enabled=amp_dtype is not None):
        pred = model(torch.cat((x, us)))
        px, pu = pred[:nl], pred[nl:]
        loss_x = F.cross_entropy(px, y)
        loss_u = (F.cross_entropy(pu, lab, reduction="none") * (conf >= thr)).mean()
        loss = (loss_x + loss_u) / 2.0
    opt.zero_grad()
    if scaler is not None:
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
    else:
        loss.backward()
        opt.step()
    torch.cuda.synchronize()
    return loss.item(), loss_x.item(), loss_u.item()


def run(leads, L, Bl, Bu, base, stem, steps=30, warmup=8, modes=("fp32_tf32", "bf16_autocast", "fp16_autocast_gradscaler"), seed=0):
    """{mode: {"samples_per_s", "ms_per_step"}} for the FixMatch step at the given shape on the current CUDA device."""
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.backends.cudnn.benchmark = True      # as the reference sets it (fixmatch.py:207)
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(Bl, leads, L, generator=g).to(dev)
    y = torch.randint(0, 4, (Bl, L), generator=g).to(dev)
    uw = torch.randn(Bu, leads, L, generator=g).to(dev)
    us = (uw + 0.5 * torch.randn(Bu, leads, L, generator=g).to(dev))
    out = {}
    for mode in modes:
        torch.manual_seed(seed)
        model = SegNet(leads, base, stem).to(dev)
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.999), weight_decay=0.05)
        amp = {"fp32_tf32": None, "bf16_autocast": torch.bfloat16, "fp16_autocast_gradscaler": torch.float16}[mode]
        scaler = torch.amp.GradScaler("cuda") if amp is torch.float16 else None
        for _ in range(warmup):
            fixmatch_step(model, opt, scaler, amp, x, y, uw, us, 0.8)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(steps):
            fixmatch_step(model, opt, scaler, amp, x, y, uw, us, 0.8)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[mode] = {"samples_per_s": round((Bl + Bu) / (ms / 1e3), 1), "ms_per_step": round(ms, 3),
                     "wall_ms_per_step": round((time.time() - t0) / steps * 1e3, 3)}
        del model, opt
    out["what"] = ("the same FixMatch step as torch.nn modules + autograd + torch.optim.AdamW, eager, cuDNN "
                   f"{torch.backends.cudnn.version()}, cudnn.benchmark on, torch {torch.__version__}; device-resident batch; "
                   "per-step loss .item() reads + synchronize as in the reference's loop; CUDA events")
    return out


if __name__ == "__main__":
    import json
    print(json.dumps(run(1, 2500, 16, 16, 64, 64)))
