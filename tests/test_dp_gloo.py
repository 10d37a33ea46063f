"""world_size-2 gloo tests (CPU) for the data-parallel host logic and the DP-equivalence claim:
per-rank gradients of equal shards, all-reduced and averaged, with BatchNorm statistics exchanged
as (sum x, sum x^2) totals and a count multiplier, equal the single-process gradients on the
concatenated batch (reference DDP + SyncBN, fixmatch.py:288-296)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import O, REPO, TINY_ARCH, TRAIN_CFG, batches, rel_err


class _SyncSum(torch.autograd.Function):
    """all-reduce(sum) in forward and backward == what the SyncBN statistic exchange does."""

    @staticmethod
    def forward(ctx, t):
        t = t.clone()
        dist.all_reduce(t)
        return t

    @staticmethod
    def backward(ctx, g):
        g = g.clone()
        dist.all_reduce(g)
        return g


def _sync_batchnorm(x, sd, pre, world):
    n = x.shape[0] * x.shape[2] * world
    s1 = _SyncSum.apply(x.sum(dim=(0, 2)))
    s2 = _SyncSum.apply((x * x).sum(dim=(0, 2)))
    mean = s1 / n
    var = s2 / n - mean * mean
    inv = 1.0 / torch.sqrt(var + 1e-5)
    return (x - mean[None, :, None]) * (inv * sd[pre + ".weight"])[None, :, None] + sd[pre + ".bias"][None, :, None]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(REPO, "semi-seg-ecg_b200", "src"))
    import utils.misc as misc
    assert misc.get_rank() == rank and misc.get_world_size() == world and misc.is_main_process() == (rank == 0)
    assert misc.all_reduce_mean(float(rank)) == pytest.approx(0.5)
    sv = misc.SmoothedValue()
    sv.update(float(rank + 1))
    sv.synchronize_between_processes()
    assert sv.count == 2 and sv.total == pytest.approx(3.0)

    torch.manual_seed(0)   # identical init on both ranks (what the rank-0 broadcast guarantees)
    import models.backbones  # noqa: F401
    from algorithms.base import init_model_from_cfg
    from helpers import model_cfg
    sd = {k: (v.double() if v.is_floating_point() else v) for k, v in init_model_from_cfg(model_cfg(2, 8, 8, 16, 0.0)).state_dict().items()}
    (lab, unl), = batches(5, 1, 4, 4, 2, 300)
    thr = 0.3

    def grads_for(xl, yl, xw, xs, bn_fn):
        leaf = {k: (v.clone().requires_grad_(True) if k in O.param_names(TINY_ARCH) else v) for k, v in sd.items()}
        orig = O.batchnorm
        if bn_fn is not None:
            O.batchnorm = lambda x, s, pre, train, nb, **kw: bn_fn(x, s, pre) if train else orig(x, s, pre, train, nb)
        try:
            with torch.no_grad():
                pw = O.forward(sd, xw.double(), TINY_ARCH, False)["seg_logits"]
                conf, label, mask = O.pseudo_label(pw, thr)
            out = O.forward(leaf, torch.cat((xl, xs)).double(), TINY_ARCH, True)["seg_logits"]
            nl = xl.shape[0]
            loss = (O.ce_hard(out[:nl], yl) + O.ce_masked(out[nl:], label, mask)) / 2
            loss.backward()
        finally:
            O.batchnorm = orig
        return {k: leaf[k].grad for k in O.param_names(TINY_ARCH)}, float(loss.detach())

    half = slice(rank * 2, rank * 2 + 2)
    g_local, loss_local = grads_for(lab["ecg"][half], lab["target"][half], unl["ecg"][half], unl["ecg_aug"][half],
                                    lambda x, s, pre: _sync_batchnorm(x, s, pre, world))
    flat = torch.cat([g.flatten() for g in g_local.values()])
    dist.all_reduce(flat)
    flat /= world
    lt = torch.tensor([loss_local], dtype=torch.float64)
    dist.all_reduce(lt)
    if rank == 0:
        g_full, loss_full = grads_for(lab["ecg"], lab["target"], unl["ecg"], unl["ecg_aug"], None)
        ref = torch.cat([g.flatten() for g in g_full.values()])
        q.put((rel_err(flat, ref), abs(float(lt) / world - loss_full)))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_equivalence_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    gerr, lerr = q.get(timeout=10)
    assert gerr < 1e-9 and lerr < 1e-12, (gerr, lerr)


def _eval_worker(rank, world, port, q):
    """semiseg_b200.evaluate.aggregate_eval on per-rank shards == the oracle's torchmetrics restatement fed the global
    batches (the reference gathers the global batch before each metric update, base.py:207-217)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(REPO, "semi-seg-ecg_b200", "src"))
    from oracle.eval_oracle import MeanIoU
    from semiseg_b200.evaluate import aggregate_eval
    rng = np.random.RandomState(3)
    ncls, L = 4, 40
    metric = MeanIoU(ncls)
    per_batch, tot, cnt = [], 0.0, 0
    for n_global in (6, 6, 2):                      # global batch sizes; every rank holds every second sample
        pred = rng.randint(0, 3, (n_global, L))     # class 3 never predicted
        tgt = rng.randint(0, ncls, (n_global, L))
        ce = rng.rand(n_global, L)                  # per-position CE terms
        oh = lambda a: torch.nn.functional.one_hot(torch.from_numpy(a), ncls).movedim(-1, 1)   # noqa: E731
        metric.update(oh(pred), oh(tgt))
        tot += float(ce.mean()) * n_global          # meters['loss'].update(batch mean, n)
        cnt += n_global
        mine = slice(rank, n_global, world)
        counts = torch.tensor([[[int(((p_ == c) & (t_ == c)).sum()), int((p_ == c).sum()), int((t_ == c).sum())]
                                for c in range(ncls)] for p_, t_ in zip(pred[mine], tgt[mine])], dtype=torch.int32)
        sums = torch.tensor([float(ce[mine].sum()), float(ce[mine].size)], dtype=torch.float64)
        per_batch.append((sums, counts, counts.shape[0]))
    stats, metrics = aggregate_eval(per_batch)
    q.put((rank, abs(stats["loss"] - tot / cnt), abs(metrics["MeanIoU"] - float(metric.compute()))))
    dist.barrier()
    dist.destroy_process_group()


def test_eval_aggregation_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_eval_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    for _ in range(2):
        rank, lerr, merr = q.get(timeout=10)
        assert lerr < 1e-12 and merr < 1e-12, (rank, lerr, merr)


def test_bench_rank_sharding_is_disjoint():
    """bench.py gives every rank its own synthetic shard (seed + rank), like fixmatch.py:204-206."""
    sys.path.insert(0, REPO)
    import bench
    a = bench.make_host_batch(0, 0, 4, 4, 1, 256)
    b = bench.make_host_batch(0, 1, 4, 4, 1, 256)
    assert not np.array_equal(a[0]["ecg"], b[0]["ecg"])
    assert np.array_equal(a[0]["ecg"], bench.make_host_batch(0, 0, 4, 4, 1, 256)[0]["ecg"])


def test_bench_untimed_steps_do_not_depend_on_the_rank():
    """Every rank must run the same number of steps around the timed region (each step holds the gradient all-reduce and the
    SyncBN exchange): bench.py's pre-roll count at N > 1 comes from the workload alone -- no clock, no rank, no environment --
    and the source has no other wall-clock-bounded loop around engine steps (the N = 4 hang of profiles/r2e_multi_gpu.md)."""
    import ast
    import inspect
    import bench
    for w in bench.WORKLOADS:
        n = bench.preroll_steps(w, 0.4)
        assert 20 <= n <= 1000 and n == bench.preroll_steps(w, 0.4)
    assert bench.preroll_steps(bench.DEFAULT_WORKLOAD, 0.4) == 400
    assert set(inspect.signature(bench.preroll_steps).parameters) == {"workload", "preroll_s"}
    body = ast.parse(inspect.getsource(bench.preroll_steps)).body[0]
    names = {n.id for n in ast.walk(body) if isinstance(n, ast.Name)} | {n.attr for n in ast.walk(body) if isinstance(n, ast.Attribute)}
    assert not names & {"time", "rank", "environ", "os", "dist", "torch"}, names
    # while-loops of bench.main that call the step: the only clock-bounded one is guarded by `n_fixed is not None`
    tree = ast.parse(inspect.getsource(bench))
    for node in ast.walk(tree):
        if isinstance(node, ast.While) and "step_from" in ast.unparse(node):
            assert "n_fixed" in ast.unparse(node.test), ast.unparse(node.test)

