"""Op-level parity of the CUDA kernels (through the C ABI) against the oracle / plain torch fp64.

FP32 kernels: relative L2 error <= 1e-5 (north_star FP32 tolerance); BF16 kernels: <= 2e-2.
Pseudo-label outputs: bit-exact against torch's own CUDA softmax/argmax on identical logits.
"""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import O, rel_err

pytestmark = pytest.mark.gpu

from semiseg_b200 import _lib  # noqa: E402
from semiseg_b200._lib import BN, Geom, StepParams, call  # noqa: E402

DEV = "cuda"
TOL = {_lib.F32: 1e-5, _lib.BF16: 2e-2}
TDT = {_lib.F32: torch.float32, _lib.BF16: torch.bfloat16}


@pytest.fixture(autouse=True, scope="module")
def _prepare_device():
    _lib.prepare()      # ssb_prepare(): device check + opt-in shared-memory sizes


def st():
    return torch.cuda.current_stream().cuda_stream


def to_flat(x_ncl, pitch, dtype):
    B, Cc, L = x_ncl.shape
    buf = torch.zeros(B * pitch, Cc, dtype=TDT[dtype], device=DEV)
    buf.view(B, pitch, Cc)[:, 1:1 + L, :] = x_ncl.permute(0, 2, 1).to(TDT[dtype])
    return buf


def from_flat(buf, B, pitch, L):
    return buf.view(B, pitch, -1)[:, 1:1 + L, :].permute(0, 2, 1).double()


def halo_is_zero(buf, B, pitch, L):
    v = buf.view(B, pitch, -1).float()
    return float(v[:, 0].abs().max()) == 0.0 and float(v[:, 1 + L:].abs().max()) == 0.0


def rq(x, dtype):
    """round-trip through the storage dtype (so the reference sees the same inputs)"""
    return x.to(TDT[dtype]).double()


def tap_major(w, dtype):
    """reference [Cout, Cin, k] -> the library's weight layout [k][Cin][Cout] in the storage dtype"""
    return w.permute(2, 1, 0).contiguous().to(TDT[dtype])


CONV_CASES = [(8, 16, 3, 1, 37), (16, 16, 3, 2, 37), (8, 16, 1, 2, 38), (16, 8, 1, 1, 20), (64, 64, 3, 1, 157),
              (64, 128, 3, 2, 313), (64, 128, 1, 2, 313), (128, 128, 3, 1, 79), (256, 512, 3, 2, 157), (512, 128, 3, 1, 79)]


def _algos(cin, cout):
    out = [(_lib.F32, _lib.ALGO_SIMT), (_lib.BF16, _lib.ALGO_SIMT)]
    if cin % 64 == 0 and cout % 64 == 0:
        out.append((_lib.BF16, _lib.ALGO_TCGEN05))
    return out


ALGO_CASES = [pytest.param(_lib.F32, _lib.ALGO_SIMT, id="fp32-simt"), pytest.param(_lib.BF16, _lib.ALGO_SIMT, id="bf16-simt"),
              pytest.param(_lib.BF16, _lib.ALGO_TCGEN05, id="bf16-tcgen05")]


@pytest.mark.parametrize("dtype,algo", ALGO_CASES)
@pytest.mark.parametrize("cin,cout,k,stride,L", CONV_CASES)
def test_conv_fwd_dgrad_wgrad(cin, cout, k, stride, L, dtype, algo):
    if algo == _lib.ALGO_TCGEN05 and (cin % 64 or cout % 64):
        pytest.skip("tcgen05 path needs channel counts that are multiples of 64")
    torch.manual_seed(cin * 7 + cout + k + stride)
    B = 3
    Lo = (L - 1) // stride + 1
    po = Lo + 2 + 3
    pi = stride * po
    x = torch.randn(B, cin, L, device=DEV, dtype=torch.float64)
    w = torch.randn(cout, cin, k, device=DEV, dtype=torch.float64) / (cin * k) ** 0.5
    dy = torch.randn(B, cout, Lo, device=DEV, dtype=torch.float64)
    gi, go = Geom(B, pi, L, cin), Geom(B, po, Lo, cout)
    for _ in (0,):
        xq, wq, dyq = rq(x, dtype), rq(w, dtype), rq(dy, dtype)
        xr = xq.clone().requires_grad_(True)
        wr = wq.clone().requires_grad_(True)
        yr = F.conv1d(xr, wr, None, stride=stride, padding=k // 2)
        assert yr.shape[2] == Lo
        yr.backward(dyq)
        wt = tap_major(w.float(), dtype)
        xb = to_flat(x, pi, dtype)
        yb = torch.full((B * po, cout), 7.0, dtype=TDT[dtype], device=DEV)
        call("ssb_conv1d_fwd", xb.data_ptr(), wt.data_ptr(), yb.data_ptr(), gi, go, k, stride, dtype, algo, st())
        tag = f"dtype={dtype} algo={algo}"
        assert rel_err(from_flat(yb, B, po, Lo), yr.detach()) < TOL[dtype], "fwd " + tag
        assert halo_is_zero(yb, B, po, Lo), "fwd halo " + tag
        # same conv with the BatchNorm statistics fused into the epilogue: identical output, and the sums
        # are those of the stored values (what the separate ssb_bn_stats pass computes)
        yb2 = torch.full((B * po, cout), 5.0, dtype=TDT[dtype], device=DEV)
        sums = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
        call("ssb_conv1d_fwd_stats", xb.data_ptr(), wt.data_ptr(), yb2.data_ptr(), gi, go, k, stride,
             sums.data_ptr(), dtype, algo, st())
        assert torch.equal(yb2, yb), "fwd+stats output " + tag
        ys = yb2.double()
        assert rel_err(sums[:cout], ys.sum(0)) < 5e-6 and rel_err(sums[cout:], (ys * ys).sum(0)) < 5e-6, "fused stats " + tag
        # dgrad (plain, then accumulate)
        dyb = to_flat(dy, po, dtype)
        dxb = torch.full((B * pi, cin), 3.0, dtype=TDT[dtype], device=DEV)
        call("ssb_conv1d_dgrad", dyb.data_ptr(), wt.data_ptr(), dxb.data_ptr(), gi, go, k, stride, 0, dtype, algo, st())
        assert rel_err(from_flat(dxb, B, pi, L), xr.grad) < TOL[dtype], "dgrad " + tag
        assert halo_is_zero(dxb, B, pi, L), "dgrad halo " + tag
        base = torch.randn(B, cin, L, device=DEV, dtype=torch.float64)
        dxb2 = to_flat(base, pi, dtype)
        call("ssb_conv1d_dgrad", dyb.data_ptr(), wt.data_ptr(), dxb2.data_ptr(), gi, go, k, stride, 1, dtype, algo, st())
        assert rel_err(from_flat(dxb2, B, pi, L), xr.grad + rq(base, dtype)) < TOL[dtype] * 2, "dgrad acc " + tag
        assert halo_is_zero(dxb2, B, pi, L)
        # wgrad (accumulates into fp32 [k][Cin][Cout], the weight layout)
        dw = torch.zeros(k, cin, cout, dtype=torch.float32, device=DEV)
        call("ssb_conv1d_wgrad", xb.data_ptr(), dyb.data_ptr(), dw.data_ptr(), gi, go, k, stride, dtype, algo, st())
        assert rel_err(dw.permute(2, 1, 0), wr.grad) < TOL[dtype], "wgrad " + tag
        call("ssb_conv1d_wgrad", xb.data_ptr(), dyb.data_ptr(), dw.data_ptr(), gi, go, k, stride, dtype, algo, st())
        assert rel_err(dw.permute(2, 1, 0), 2 * wr.grad) < TOL[dtype], "wgrad accumulate " + tag


def test_weight_shadow():
    torch.manual_seed(0)
    src = torch.randn(4096 + 64, device=DEV)
    dst = torch.zeros(src.numel(), dtype=torch.bfloat16, device=DEV)
    call("ssb_weight_shadow", src.data_ptr(), dst.data_ptr(), src.numel(), _lib.BF16, st())
    assert torch.equal(dst, src.to(torch.bfloat16))
    with pytest.raises(RuntimeError):
        call("ssb_weight_shadow", src.data_ptr(), dst.data_ptr(), 12, _lib.BF16, st())


def make_bn(Cn, train_count_mul=0):
    t = {"gamma": torch.rand(Cn, device=DEV) + 0.5, "beta": torch.randn(Cn, device=DEV) * 0.1,
         "rm": torch.randn(Cn, device=DEV) * 0.1, "rv": torch.rand(Cn, device=DEV) + 0.5,
         "nbt": torch.zeros(1, dtype=torch.int64, device=DEV), "sums": torch.zeros(2 * Cn, dtype=torch.float64, device=DEV),
         "mi": torch.zeros(2 * Cn, device=DEV), "bsums": torch.zeros(2 * Cn, dtype=torch.float64, device=DEV),
         "dgamma": torch.zeros(Cn, device=DEV), "dbeta": torch.zeros(Cn, device=DEV)}
    s = BN()
    s.gamma, s.beta, s.running_mean, s.running_var = t["gamma"].data_ptr(), t["beta"].data_ptr(), t["rm"].data_ptr(), t["rv"].data_ptr()
    s.num_batches_tracked, s.sums, s.mean_invstd, s.bwd_sums = t["nbt"].data_ptr(), t["sums"].data_ptr(), t["mi"].data_ptr(), t["bsums"].data_ptr()
    s.dgamma, s.dbeta, s.count_mul = t["dgamma"].data_ptr(), t["dbeta"].data_ptr(), train_count_mul
    return s, t


def ref_bn(x, t, train):
    sd = {"p.weight": t["gamma"].double(), "p.bias": t["beta"].double(), "p.running_mean": t["rm"].double(),
          "p.running_var": t["rv"].double(), "p.num_batches_tracked": t["nbt"][0].clone()}
    nb = {}
    y = O.batchnorm(x, sd, "p", train, nb)
    return y, nb


@pytest.mark.parametrize("dtype,algo", ALGO_CASES)
@pytest.mark.parametrize("cin,cout,k,stride,L,with_res,relu", [(64, 64, 3, 1, 157, True, 1), (64, 128, 1, 2, 313, False, 0),
                                                              (128, 256, 3, 2, 157, False, 1), (512, 128, 3, 1, 79, False, 1),
                                                              (16, 8, 3, 1, 37, True, 1)])
def test_conv_bn_act_eval(cin, cout, k, stride, L, with_res, relu, dtype, algo):
    """eval-mode conv + BatchNorm(running stats) [+ residual] [+ ReLU] in one launch vs the composed reference"""
    if algo == _lib.ALGO_TCGEN05 and (cin % 64 or cout % 64):
        pytest.skip("tcgen05 path needs channel counts that are multiples of 64")
    torch.manual_seed(cin + cout + k)
    B = 3
    Lo = (L - 1) // stride + 1
    po = Lo + 2 + 2
    pi = stride * po
    x = torch.randn(B, cin, L, device=DEV, dtype=torch.float64)
    w = torch.randn(cout, cin, k, device=DEV, dtype=torch.float64) / (cin * k) ** 0.5
    r = torch.randn(B, cout, Lo, device=DEV, dtype=torch.float64)
    bn, t = make_bn(cout)
    rm0, rv0 = t["rm"].clone(), t["rv"].clone()
    y_ref, _ = ref_bn(F.conv1d(rq(x, dtype), rq(w, dtype), None, stride=stride, padding=k // 2), t, False)
    if with_res:
        y_ref = y_ref + rq(r, dtype)
    if relu:
        y_ref = torch.relu(y_ref)
    gi, go = Geom(B, pi, L, cin), Geom(B, po, Lo, cout)
    xb, rb, wt = to_flat(x, pi, dtype), to_flat(r, po, dtype), tap_major(w.float(), dtype)
    yb = torch.full((B * po, cout), 7.0, dtype=TDT[dtype], device=DEV)
    call("ssb_conv1d_bn_act_fwd", xb.data_ptr(), wt.data_ptr(), yb.data_ptr(), gi, go, k, stride, C.byref(bn),
         rb.data_ptr() if with_res else None, relu, dtype, algo, st())
    assert rel_err(from_flat(yb, B, po, Lo), y_ref) < TOL[dtype]
    assert halo_is_zero(yb, B, po, Lo)
    assert torch.equal(t["rm"], rm0) and torch.equal(t["rv"], rv0) and int(t["nbt"][0]) == 0   # eval mode: buffers untouched


@pytest.mark.parametrize("dtype,algo", ALGO_CASES)
@pytest.mark.parametrize("cin,cout,k,stride,L,with_res", [(64, 64, 3, 1, 157, True), (64, 128, 3, 2, 313, False),
                                                         (128, 256, 1, 2, 157, False), (256, 256, 3, 1, 79, True),
                                                         (16, 8, 3, 1, 37, True)])
def test_conv_fwd_dual(cin, cout, k, stride, L, with_res, dtype, algo):
    """train rows (raw output + statistics) and eval rows (BN running stats [+ residual] + ReLU) of one conv in one launch"""
    if algo == _lib.ALGO_TCGEN05 and (cin % 64 or cout % 64):
        pytest.skip("tcgen05 path needs channel counts that are multiples of 64")
    torch.manual_seed(cin + cout + k + 1)
    B, Bt = 5, 3
    Lo = (L - 1) // stride + 1
    po = Lo + 2 + 1
    pi = stride * po
    x = torch.randn(B, cin, L, device=DEV, dtype=torch.float64)
    w = torch.randn(cout, cin, k, device=DEV, dtype=torch.float64) / (cin * k) ** 0.5
    r = torch.randn(B, cout, Lo, device=DEV, dtype=torch.float64)
    bn, t = make_bn(cout)
    conv = F.conv1d(rq(x, dtype), rq(w, dtype), None, stride=stride, padding=k // 2)
    y_eval_ref, _ = ref_bn(conv[Bt:], t, False)
    if with_res:
        y_eval_ref = y_eval_ref + rq(r, dtype)[Bt:]
    y_eval_ref = torch.relu(y_eval_ref)
    gi, go = Geom(B, pi, L, cin), Geom(B, po, Lo, cout)
    xb, rb, wt = to_flat(x, pi, dtype), to_flat(r, po, dtype), tap_major(w.float(), dtype)
    y_tr = torch.full((B * po, cout), 7.0, dtype=TDT[dtype], device=DEV)
    y_ev = torch.full((B * po, cout), 9.0, dtype=TDT[dtype], device=DEV)
    sums = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    call("ssb_conv1d_fwd_dual", xb.data_ptr(), wt.data_ptr(), y_tr.data_ptr(), y_ev.data_ptr(), gi, go, k, stride, Bt,
         sums.data_ptr(), C.byref(bn), rb.data_ptr() if with_res else None, 1, dtype, algo, st())
    tr = y_tr.view(B, po, cout)
    ev = y_ev.view(B, po, cout)
    assert rel_err(from_flat(y_tr, B, po, Lo)[:Bt], conv[:Bt]) < TOL[dtype]
    assert rel_err(from_flat(y_ev, B, po, Lo)[Bt:], y_eval_ref) < TOL[dtype]
    assert halo_is_zero(tr[:Bt].reshape(-1, cout), Bt, po, Lo) and halo_is_zero(ev[Bt:].reshape(-1, cout), B - Bt, po, Lo)
    assert float((tr[Bt:].float() - 7.0).abs().max()) == 0.0 and float((ev[:Bt].float() - 9.0).abs().max()) == 0.0   # other range untouched
    ys = tr[:Bt].reshape(-1, cout).double()
    assert rel_err(sums[:cout], ys.sum(0)) < 5e-6 and rel_err(sums[cout:], (ys * ys).sum(0)) < 5e-6


@pytest.mark.parametrize("dtype,algo", ALGO_CASES)
@pytest.mark.parametrize("cin,cout,k,L,acc,with_res,B", [(64, 64, 3, 157, 0, False, 3), (128, 128, 3, 79, 1, True, 3),
                                                         (256, 128, 3, 79, 1, False, 3), (512, 512, 3, 40, 0, True, 3),
                                                         (16, 8, 3, 37, 1, True, 3),
                                                         # the benchmark's batch (config 2, B = 32): every stage's k3 conv
                                                         (64, 64, 3, 625, 1, True, 32), (128, 128, 3, 313, 0, False, 32),
                                                         (256, 256, 3, 157, 1, True, 32), (512, 512, 3, 79, 0, False, 32),
                                                         (512, 128, 3, 79, 0, True, 32)])
def test_conv_dgrad_bnred(cin, cout, k, L, acc, with_res, B, dtype, algo):
    """dgrad with the BN-backward reduce of the produced gradient fused in == dgrad followed by ssb_bn_bwd_reduce"""
    if algo == _lib.ALGO_TCGEN05 and (cin % 64 or cout % 64):
        pytest.skip("tcgen05 path needs channel counts that are multiples of 64")
    if B > 3 and algo != _lib.ALGO_TCGEN05:
        pytest.skip("large cases: tensor-core path only")
    torch.manual_seed(cin + cout + L)
    pitch = L + 2 + 2
    g = Geom(B, pitch, L, cin)
    go = Geom(B, pitch, L, cout)
    w = torch.randn(cout, cin, k, device=DEV, dtype=torch.float64) / (cout * k) ** 0.5
    dy = to_flat(torch.randn(B, cout, L, device=DEV, dtype=torch.float64), pitch, dtype)
    y_act = to_flat(torch.relu(torch.randn(B, cin, L, device=DEV, dtype=torch.float64)), pitch, dtype)
    x_pre = to_flat(torch.randn(B, cin, L, device=DEV, dtype=torch.float64) * 1.3 + 0.2, pitch, dtype)
    x_res = to_flat(torch.randn(B, cin, L, device=DEV, dtype=torch.float64), pitch, dtype)
    base = to_flat(torch.randn(B, cin, L, device=DEV, dtype=torch.float64), pitch, dtype)
    wt = tap_major(w.float(), dtype)
    outs = []
    for fused in (False, True):
        bn, t = make_bn(cin)
        bnr, tr_ = make_bn(cin)
        t["mi"][:cin] = 0.2
        t["mi"][cin:] = 0.8
        tr_["mi"][:cin] = -0.1
        tr_["mi"][cin:] = 1.1
        dx = base.clone()
        if fused:
            call("ssb_conv1d_dgrad_bnred", dy.data_ptr(), wt.data_ptr(), dx.data_ptr(), g, go, k, 1, acc, y_act.data_ptr(),
                 x_pre.data_ptr(), C.byref(bn), x_res.data_ptr() if with_res else None, C.byref(bnr) if with_res else None,
                 dtype, algo, st())
        else:
            call("ssb_conv1d_dgrad", dy.data_ptr(), wt.data_ptr(), dx.data_ptr(), g, go, k, 1, acc, dtype, algo, st())
            call("ssb_bn_bwd_reduce", dx.data_ptr(), None, y_act.data_ptr(), x_pre.data_ptr(), C.byref(bn),
                 x_res.data_ptr() if with_res else None, C.byref(bnr) if with_res else None, g, dtype, st())
        torch.cuda.synchronize()
        outs.append((dx, t["bsums"].clone(), tr_["bsums"].clone()))
    (dx0, s0, r0), (dx1, s1, r1) = outs
    assert torch.equal(dx0, dx1)
    assert rel_err(s1, s0) < 5e-6
    if with_res:
        assert rel_err(r1, r0) < 5e-6


@pytest.mark.parametrize("cin,cout,k,stride,L,res_mode", [(64, 64, 3, 1, 625, 1), (64, 128, 3, 2, 625, 0), (128, 128, 3, 1, 313, 2),
                                                          (256, 256, 3, 1, 157, 1), (512, 128, 3, 1, 79, 0), (128, 256, 1, 2, 313, 0)])
def test_conv_fwd_bn_train(cin, cout, k, stride, L, res_mode):
    """train-mode conv + BN(batch stats) [+ residual | + BN(residual)] + ReLU in one launch (grid barrier, second pass over
    the accumulator) == ssb_conv1d_fwd_stats followed by ssb_bn_act_fwd, including the BatchNorm state it leaves behind"""
    dtype, algo = _lib.BF16, _lib.ALGO_TCGEN05
    torch.manual_seed(cin + cout + L)
    B = 16
    Lo = (L - 1) // stride + 1
    po = Lo + 2 + 1
    pi = stride * po
    gi, go = Geom(B, pi, L, cin), Geom(B, po, Lo, cout)
    assert _lib.load().ssb_conv1d_fwd_bn_train_fits(gi, go, k, stride, dtype, algo) == 1
    x = to_flat(torch.randn(B, cin, L, device=DEV, dtype=torch.float64), pi, dtype)
    w = tap_major((torch.randn(cout, cin, k, device=DEV) / (cin * k) ** 0.5), dtype)
    r = to_flat(torch.randn(B, cout, Lo, device=DEV, dtype=torch.float64), po, dtype)
    res = []
    for fused in (False, True):
        torch.manual_seed(99)
        bn, t = make_bn(cout)
        bnr, tr_ = make_bn(cout)
        if res_mode == 2:
            call("ssb_bn_stats", r.data_ptr(), go, tr_["sums"].data_ptr(), dtype, st())
        y_raw = torch.full((B * po, cout), 7.0, dtype=TDT[dtype], device=DEV)
        y_act = torch.full((B * po, cout), 9.0, dtype=TDT[dtype], device=DEV)
        rp = r.data_ptr() if res_mode else None
        rb = C.byref(bnr) if res_mode == 2 else None
        if fused:
            bar = torch.zeros(1, dtype=torch.int32, device=DEV)
            call("ssb_conv1d_fwd_bn_train", x.data_ptr(), w.data_ptr(), y_raw.data_ptr(), y_act.data_ptr(), gi, go, k, stride,
                 C.byref(bn), rp, rb, 1, bar.data_ptr(), dtype, algo, st())
        else:
            call("ssb_conv1d_fwd_stats", x.data_ptr(), w.data_ptr(), y_raw.data_ptr(), gi, go, k, stride, t["sums"].data_ptr(),
                 dtype, algo, st())
            call("ssb_bn_act_fwd", y_raw.data_ptr(), C.byref(bn), rp, rb, y_act.data_ptr(), go, 1, 1, dtype, st())
        torch.cuda.synchronize()
        res.append((y_raw.clone(), y_act.float(), t["rm"].clone(), t["rv"].clone(), t["mi"].clone(), int(t["nbt"][0]),
                    tr_["rm"].clone(), tr_["mi"].clone()))
    a_, b_ = res
    assert torch.equal(a_[0], b_[0])
    assert rel_err(b_[1], a_[1]) < 1e-2 and float((b_[1] - a_[1]).abs().max()) < 0.05 and halo_is_zero(b_[1], B, po, Lo)
    assert rel_err(b_[2], a_[2]) < 1e-5 and rel_err(b_[3], a_[3]) < 1e-5 and rel_err(b_[4], a_[4]) < 1e-5 and a_[5] == b_[5] == 1
    if res_mode == 2:
        assert rel_err(b_[6], a_[6]) < 1e-5 and rel_err(b_[7], a_[7]) < 1e-5
    # a conv with more tiles than co-resident CTAs is refused by the query
    assert _lib.load().ssb_conv1d_fwd_bn_train_fits(Geom(512, 1296, 1250, 128), Geom(512, 1296, 1250, 128), 3, 1, dtype, algo) == 0


@pytest.mark.parametrize("dtype", [_lib.F32, _lib.BF16])
@pytest.mark.parametrize("cluster", ["16", "8", "0"])
@pytest.mark.parametrize("Cn,res_mode,L,B", [(64, 0, 625, 16), (128, 1, 313, 16), (512, 2, 79, 16), (24, 1, 50, 16), (8, 2, 33, 16),
                                             (64, 1, 625, 32), (128, 2, 313, 32), (256, 0, 157, 32), (512, 1, 79, 32), (1024, 2, 157, 32),
                                             (128, 0, 1250, 64)])
def test_bn_bwd_fused(dtype, Cn, res_mode, L, B, cluster, monkeypatch):
    """both BatchNorm-backward passes in one launch == ssb_bn_bwd_reduce followed by ssb_bn_bwd_apply.  cluster 16 / 8: the
    thread-block-cluster kernel (register-resident rows, partial sums exchanged through distributed shared memory) where
    the tensor fits it; 0: the grid-barrier kernel.  B = 32: the benchmark's batch (config 2 / config 5 at 16+16)."""
    monkeypatch.setenv("SSB_BN_CLUSTER", cluster)
    torch.manual_seed(Cn + L)
    pitch = L + 3
    g = Geom(B, pitch, L, Cn)
    gr = to_flat(torch.randn(B, Cn, L, device=DEV, dtype=torch.float64), pitch, dtype)
    yy = to_flat(torch.relu(torch.randn(B, Cn, L, device=DEV, dtype=torch.float64)), pitch, dtype)
    xx = to_flat(torch.randn(B, Cn, L, device=DEV, dtype=torch.float64) + 0.3, pitch, dtype)
    xr = to_flat(torch.randn(B, Cn, L, device=DEV, dtype=torch.float64), pitch, dtype)
    res = []
    for fused in (False, True):
        torch.manual_seed(1234)          # same gamma / beta in both runs
        bn, t = make_bn(Cn)
        bnr, tr_ = make_bn(Cn)
        for d in (t, tr_):
            d["mi"][:Cn] = 0.1
            d["mi"][Cn:] = 0.9
        dx = torch.full((B * pitch, Cn), 3.0, dtype=TDT[dtype], device=DEV)
        dxr = torch.full((B * pitch, Cn), 3.0, dtype=TDT[dtype], device=DEV)
        gid = torch.full((B * pitch, Cn), 3.0, dtype=TDT[dtype], device=DEV)
        a_res = (xr.data_ptr(), C.byref(bnr), dxr.data_ptr(), None) if res_mode == 2 else (None, None, None, gid.data_ptr() if res_mode == 1 else None)
        if fused:
            bar = torch.zeros(1, dtype=torch.int32, device=DEV)
            rep = torch.zeros(2, 8, 2 * Cn + 64, dtype=torch.float64, device=DEV)
            call("ssb_bn_bwd_fused", gr.data_ptr(), yy.data_ptr(), xx.data_ptr(), C.byref(bn), dx.data_ptr(), a_res[0], a_res[1],
                 a_res[2], a_res[3], g, bar.data_ptr(), rep[0].data_ptr(), rep[1].data_ptr() if res_mode == 2 else None,
                 2 * Cn + 64, dtype, st())
        else:
            call("ssb_bn_bwd_reduce", gr.data_ptr(), None, yy.data_ptr(), xx.data_ptr(), C.byref(bn), a_res[0], a_res[1], g, dtype, st())
            call("ssb_bn_bwd_apply", gr.data_ptr(), None, yy.data_ptr(), xx.data_ptr(), C.byref(bn), dx.data_ptr(), a_res[0], a_res[1],
                 a_res[2], a_res[3], g, dtype, st())
        torch.cuda.synchronize()
        res.append((dx.float(), dxr.float(), gid.float(), t["dgamma"].clone(), t["dbeta"].clone(), tr_["dgamma"].clone()))
    tol = 1e-5 if dtype == _lib.F32 else 1e-2
    assert rel_err(res[1][0], res[0][0]) < tol and halo_is_zero(res[1][0], B, pitch, L)
    assert rel_err(res[1][3], res[0][3]) < 1e-5 and rel_err(res[1][4], res[0][4]) < 1e-5
    if res_mode == 2:
        assert rel_err(res[1][1], res[0][1]) < tol and rel_err(res[1][5], res[0][5]) < 1e-5
    if res_mode == 1:
        assert torch.equal(res[1][2], res[0][2])
    assert _lib.load().ssb_bn_bwd_fused_fits(Geom(64, 1300, 1250, 64), res_mode, 1, dtype) == 1   # any size: the grid is capped


@pytest.mark.parametrize("dtype", [_lib.F32, _lib.BF16])
@pytest.mark.parametrize("Cn,res_mode", [(8, 0), (64, 1), (128, 2), (24, 0)])
def test_bn_forward_backward(dtype, Cn, res_mode):
    torch.manual_seed(Cn + res_mode)
    B, L, pitch = 4, 53, 60
    g = Geom(B, pitch, L, Cn)
    x = torch.randn(B, Cn, L, device=DEV, dtype=torch.float64) * 1.5 + 0.3
    r = torch.randn(B, Cn, L, device=DEV, dtype=torch.float64)
    gout = torch.randn(B, Cn, L, device=DEV, dtype=torch.float64)
    xq, rq_, gq = rq(x, dtype), rq(r, dtype), rq(gout, dtype)
    bn, t = make_bn(Cn)
    bnr, tr_ = make_bn(Cn)
    # ---- reference (train mode) ----
    xr = xq.clone().requires_grad_(True)
    rr = rq_.clone().requires_grad_(True)
    rm0, rv0 = t["rm"].clone(), t["rv"].clone()
    y_ref, nb = ref_bn(xr, t, True)
    if res_mode == 1:
        y_ref = y_ref + rr
    elif res_mode == 2:
        yr2, nbr = ref_bn(rr, tr_, True)
        y_ref = y_ref + yr2
    y_ref = torch.relu(y_ref)
    y_ref.backward(gq)
    # ---- CUDA ----
    xb, rb = to_flat(x, pitch, dtype), to_flat(r, pitch, dtype)
    yb = torch.full((B * pitch, Cn), 5.0, dtype=TDT[dtype], device=DEV)
    call("ssb_bn_stats", xb.data_ptr(), g, t["sums"].data_ptr(), dtype, st())
    if res_mode == 2:
        call("ssb_bn_stats", rb.data_ptr(), g, tr_["sums"].data_ptr(), dtype, st())
    call("ssb_bn_act_fwd", xb.data_ptr(), C.byref(bn), rb.data_ptr() if res_mode else None,
         C.byref(bnr) if res_mode == 2 else None, yb.data_ptr(), g, 1, 1, dtype, st())
    tol = TOL[dtype]
    assert rel_err(from_flat(yb, B, pitch, L), y_ref.detach()) < tol
    assert halo_is_zero(yb, B, pitch, L)
    assert rel_err(t["rm"], nb["p.running_mean"]) < 1e-5 and rel_err(t["rv"], nb["p.running_var"]) < 1e-5
    assert int(t["nbt"][0]) == 1
    # backward: reduce + apply
    gb = to_flat(gout, pitch, dtype)
    dxb = torch.full((B * pitch, Cn), 5.0, dtype=TDT[dtype], device=DEV)
    dxr = torch.full((B * pitch, Cn), 5.0, dtype=TDT[dtype], device=DEV)
    gid = torch.full((B * pitch, Cn), 5.0, dtype=TDT[dtype], device=DEV)
    call("ssb_bn_bwd_reduce", gb.data_ptr(), None, yb.data_ptr(), xb.data_ptr(), C.byref(bn),
         rb.data_ptr() if res_mode == 2 else None, C.byref(bnr) if res_mode == 2 else None, g, dtype, st())
    call("ssb_bn_bwd_apply", gb.data_ptr(), None, yb.data_ptr(), xb.data_ptr(), C.byref(bn), dxb.data_ptr(),
         rb.data_ptr() if res_mode == 2 else None, C.byref(bnr) if res_mode == 2 else None,
         dxr.data_ptr() if res_mode == 2 else None, gid.data_ptr() if res_mode == 1 else None, g, dtype, st())
    # the CUDA relu mask comes from the stored (rounded) output; compare against the same rule
    btol = tol * 3
    assert rel_err(from_flat(dxb, B, pitch, L), xr.grad) < btol
    assert halo_is_zero(dxb, B, pitch, L)
    if res_mode == 1:
        assert rel_err(from_flat(gid, B, pitch, L), rr.grad) < btol
    if res_mode == 2:
        assert rel_err(from_flat(dxr, B, pitch, L), rr.grad) < btol
    # dgamma / dbeta
    xhat = (xq - xq.mean(dim=(0, 2), keepdim=True)) / torch.sqrt(xq.var(dim=(0, 2), unbiased=False, keepdim=True) + 1e-5)
    gm = gq * (y_ref.detach() > 0)
    assert rel_err(t["dgamma"], (gm * xhat).sum(dim=(0, 2))) < btol
    assert rel_err(t["dbeta"], gm.sum(dim=(0, 2))) < btol
    # ---- eval mode uses running statistics ----
    t["rm"].copy_(rm0)
    t["rv"].copy_(rv0)
    y_eval, _ = ref_bn(xq, t, False)
    call("ssb_bn_act_fwd", xb.data_ptr(), C.byref(bn), None, None, yb.data_ptr(), g, 0, 0, dtype, st())
    assert rel_err(from_flat(yb, B, pitch, L), y_eval) < tol


@pytest.mark.parametrize("dtype", [_lib.F32, _lib.BF16])
@pytest.mark.parametrize("Cl,Cs,L", [(1, 64, 250), (2, 8, 301), (12, 128, 200)])
def test_stem(dtype, Cl, Cs, L):
    torch.manual_seed(Cl + Cs)
    B = 3
    L0 = (L - 1) // 2 + 1
    Lp = (L0 - 1) // 2 + 1
    pp = Lp + 2 + 1
    p0 = 2 * pp
    g0, gp = Geom(B, p0, L0, Cs), Geom(B, pp, Lp, Cs)
    x = torch.randn(B, Cl, L, device=DEV)
    w = torch.randn(Cs, Cl, 7, device=DEV) / (7 * Cl) ** 0.5
    gpool = torch.randn(B, Cs, Lp, device=DEV, dtype=torch.float64)
    bn, t = make_bn(Cs)
    # reference
    wr = w.double().clone().requires_grad_(True)
    c0r = F.conv1d(x.double(), wr, None, stride=2, padding=3)
    c0q = c0r if dtype == _lib.F32 else (c0r + (rq(c0r.detach(), dtype) - c0r.detach()))  # see stored rounding
    a0, nb = ref_bn(c0q, t, True)
    pr = O.maxpool_k3s2p1(torch.relu(a0))
    pr.backward(rq(gpool, dtype))
    # CUDA
    c0 = torch.full((B * p0, Cs), 9.0, dtype=TDT[dtype], device=DEV)
    call("ssb_stem_conv_fwd", x.data_ptr(), w.data_ptr(), c0.data_ptr(), Cl, L, g0, dtype, st())
    assert rel_err(from_flat(c0, B, p0, L0), c0r.detach()) < TOL[dtype]
    assert halo_is_zero(c0, B, p0, L0)
    call("ssb_bn_stats", c0.data_ptr(), g0, t["sums"].data_ptr(), dtype, st())
    pb = torch.full((B * pp, Cs), 9.0, dtype=TDT[dtype], device=DEV)
    arg = torch.full((B * pp, Cs), 9, dtype=torch.uint8, device=DEV)
    call("ssb_stem_bn_relu_pool_fwd", c0.data_ptr(), C.byref(bn), pb.data_ptr(), arg.data_ptr(), g0, gp, 1, dtype, st())
    assert rel_err(from_flat(pb, B, pp, Lp), pr.detach()) < TOL[dtype]
    assert halo_is_zero(pb, B, pp, Lp)
    gpb = to_flat(gpool, pp, dtype)
    dc0 = torch.full((B * p0, Cs), 9.0, dtype=TDT[dtype], device=DEV)
    call("ssb_stem_bwd_reduce", gpb.data_ptr(), c0.data_ptr(), arg.data_ptr(), C.byref(bn), g0, gp, dtype, st())
    call("ssb_stem_bwd_apply", gpb.data_ptr(), c0.data_ptr(), arg.data_ptr(), C.byref(bn), dc0.data_ptr(), g0, gp, dtype, st())
    assert halo_is_zero(dc0, B, p0, L0)
    dw = torch.zeros(Cs, Cl, 7, device=DEV)
    call("ssb_stem_conv_wgrad", x.data_ptr(), dc0.data_ptr(), dw.data_ptr(), Cl, L, g0, dtype, st())
    assert rel_err(dw, wr.grad) < TOL[dtype] * 3


@pytest.mark.parametrize("dtype", [_lib.F32, _lib.BF16])
@pytest.mark.parametrize("p", [0.0, 0.1])
def test_head_cls(dtype, p):
    torch.manual_seed(5)
    B, L, Cn, ncls, pitch = 3, 19, 128, 4, 23
    g = Geom(B, pitch, L, Cn)
    a = torch.relu(torch.randn(B, Cn, L, device=DEV, dtype=torch.float64))
    w = torch.randn(ncls, Cn, device=DEV) * 0.1
    bias = torch.randn(ncls, device=DEV) * 0.1
    mask = (torch.rand(B, L, Cn, device=DEV) >= p).to(torch.uint8)
    dlow = torch.randn(B, L, ncls, device=DEV)
    ar = rq(a, dtype).clone().requires_grad_(True)
    wr = w.double().clone().requires_grad_(True)
    br = bias.double().clone().requires_grad_(True)
    h = ar * mask.permute(0, 2, 1).double() / (1.0 - p) if p > 0 else ar
    low_r = F.conv1d(h, wr[:, :, None], br)
    low_r.backward(dlow.permute(0, 2, 1).double())
    ab = to_flat(a, pitch, dtype)
    low = torch.zeros(B, L, ncls, device=DEV)
    mp = mask.data_ptr() if p > 0 else None
    call("ssb_head_cls_fwd", ab.data_ptr(), w.data_ptr(), bias.data_ptr(), low.data_ptr(), g, ncls, p, mp, None, dtype, st())
    assert rel_err(low.permute(0, 2, 1), low_r.detach()) < 1e-5
    da = torch.full((B * pitch, Cn), 2.0, dtype=TDT[dtype], device=DEV)
    dw = torch.zeros(ncls, Cn, device=DEV)
    db = torch.zeros(ncls, device=DEV)
    call("ssb_head_cls_bwd", dlow.data_ptr(), ab.data_ptr(), w.data_ptr(), da.data_ptr(), dw.data_ptr(), db.data_ptr(), g,
         ncls, p, mp, None, dtype, st())
    assert rel_err(from_flat(da, B, pitch, L), ar.grad) < TOL[dtype]
    assert halo_is_zero(da, B, pitch, L)
    assert rel_err(dw, wr.grad) < 1e-5 and rel_err(db, br.grad) < 1e-5


def test_dropout_rng_statistics():
    """Counter-based dropout: keep-rate ~ 1-p, fwd and bwd use the same mask, masks differ per step."""
    B, L, Cn, ncls, pitch, p = 8, 79, 128, 4, 81, 0.1
    g = Geom(B, pitch, L, Cn)
    ab = to_flat(torch.ones(B, Cn, L, device=DEV, dtype=torch.float64), pitch, _lib.F32)
    w = torch.ones(ncls, Cn, device=DEV)
    bias = torch.zeros(ncls, device=DEV)
    lows = []
    for step in (0, 1):
        sp = StepParams()
        sp.rng_seed, sp.rng_step = 123, step
        spd = torch.frombuffer(bytearray(bytes(sp)), dtype=torch.uint8).to(DEV)
        low = torch.zeros(B, L, ncls, device=DEV)
        call("ssb_head_cls_fwd", ab.data_ptr(), w.data_ptr(), bias.data_ptr(), low.data_ptr(), g, ncls, p, None, spd.data_ptr(), _lib.F32, st())
        kept = low[..., 0] * (1 - p)     # number of kept channels per row
        assert abs(float(kept.mean()) / Cn - (1 - p)) < 0.01
        da = torch.zeros(B * pitch, Cn, device=DEV)
        dw = torch.zeros(ncls, Cn, device=DEV)
        db = torch.zeros(ncls, device=DEV)
        dl = torch.zeros(B, L, ncls, device=DEV)
        dl[..., 0] = 1.0
        call("ssb_head_cls_bwd", dl.data_ptr(), ab.data_ptr(), w.data_ptr(), da.data_ptr(), dw.data_ptr(), db.data_ptr(), g, ncls, p,
             None, spd.data_ptr(), _lib.F32, st())
        kept_b = (from_flat(da, B, pitch, L) > 0).sum(dim=1).float()
        assert torch.equal(kept_b, torch.round(kept).float())
        lows.append(low)
    assert not torch.equal(lows[0], lows[1])


@pytest.mark.parametrize("Lin,Lout", [(79, 2500), (157, 5000), (10, 300), (5, 5)])
def test_upsample(Lin, Lout):
    torch.manual_seed(1)
    B, ncls = 3, 4
    low = torch.randn(B, Lin, ncls, device=DEV)
    out = torch.zeros(B, ncls, Lout, device=DEV)
    call("ssb_upsample_fwd", low.data_ptr(), out.data_ptr(), B, Lin, Lout, ncls, 0, st())
    lr = low.permute(0, 2, 1).contiguous().requires_grad_(True)
    ref = F.interpolate(lr, size=Lout, mode="linear", align_corners=False)
    assert float((out - ref.detach()).abs().max()) < 1e-5
    assert rel_err(out, O.linear_upsample(low.permute(0, 2, 1).double(), Lout)) < 2e-5   # fp32 index arithmetic, like ATen
    dout = torch.randn(B, ncls, Lout, device=DEV)
    ref.backward(dout)
    dlow = torch.zeros(B, Lin, ncls, device=DEV)
    call("ssb_upsample_bwd", dout.data_ptr(), dlow.data_ptr(), B, Lin, Lout, ncls, 0, st())
    assert rel_err(dlow.permute(0, 2, 1), lr.grad) < 1e-5


def _adversarial_logits(U, L):
    torch.manual_seed(3)
    z = torch.randn(U, 4, L, device=DEV)
    z[0] *= 5
    z[1] *= 20
    z[2, :, :50] = 0.0                       # exact ties -> first index
    z[2, 1, 50:100] = z[2, 3, 50:100]        # pairwise ties
    # confidences within a few ulp of fp32(0.8): logits (a, 0, 0, 0) with softmax max = 0.8 -> a = ln(12)
    a = float(np.log(12.0))
    z[3, :, :] = 0.0
    z[3, 0, :] = a + (torch.arange(L, device=DEV) - L // 2).float() * 1e-7
    return z


def test_pseudo_label_bit_exact():
    U, L = 6, 2500
    z = _adversarial_logits(U, L)
    thr = 0.8
    conf = torch.zeros(U, L, device=DEV)
    label = torch.zeros(U, L, dtype=torch.int64, device=DEV)
    mask = torch.zeros(U, L, dtype=torch.uint8, device=DEV)
    call("ssb_pseudo_label", z.data_ptr(), thr, conf.data_ptr(), label.data_ptr(), mask.data_ptr(), U, 4, L, st())
    rc, rl, rm = O.pseudo_label(z, thr)      # torch's own CUDA softmax / argmax / compare
    assert torch.equal(label, rl)
    assert torch.equal(conf.view(torch.int32), rc.view(torch.int32))      # bit pattern
    assert torch.equal(mask.bool(), rm)
    assert 0 < int(rm[3].sum()) < L           # the 1-ulp band straddles the threshold
    # special values
    z2 = torch.randn(2, 4, 64, device=DEV)
    z2[0, 1, 3] = float("inf")
    z2[0, 2, 5] = float("-inf")
    z2[1, 0, 7] = float("nan")
    c2 = torch.zeros(2, 64, device=DEV)
    l2 = torch.zeros(2, 64, dtype=torch.int64, device=DEV)
    m2 = torch.zeros(2, 64, dtype=torch.uint8, device=DEV)
    call("ssb_pseudo_label", z2.data_ptr(), thr, c2.data_ptr(), l2.data_ptr(), m2.data_ptr(), 2, 4, 64, st())
    rc, rl, rm = O.pseudo_label(z2, thr)
    assert torch.equal(l2, rl) and torch.equal(m2.bool(), rm)
    assert torch.equal(torch.isnan(c2), torch.isnan(rc))


@pytest.mark.parametrize("mode", [_lib.LOSS_SUP, _lib.LOSS_FIXMATCH, _lib.LOSS_SOFT, _lib.LOSS_SOFT_MASKED])
@pytest.mark.parametrize("Lin,L", [(79, 2500), (10, 300), (157, 5000)])
def test_semi_loss(mode, Lin, L):
    torch.manual_seed(Lin + mode)
    Bl, Bu, ncls = 3, (0 if mode == _lib.LOSS_SUP else 4), 4
    S = Bl + Bu
    low_s = torch.randn(S, Lin, ncls, device=DEV) * 2
    low_t = torch.randn(max(Bu, 1), Lin, ncls, device=DEV) * 3
    y = torch.randint(0, ncls, (Bl, L), device=DEV)
    thr = 0.6
    # oracle (fp64) on the upsampled logits, autograd back to the low-res logits
    ls = low_s.double().permute(0, 2, 1).contiguous().requires_grad_(True)
    zs = O.linear_upsample(ls, L)
    loss_x = O.ce_hard(zs[:Bl], y)
    if mode == _lib.LOSS_SUP:
        loss, loss_u, mratio = loss_x, None, None
    else:
        zt = O.linear_upsample(low_t.permute(0, 2, 1).contiguous(), L)   # fp32, like the CUDA path
        if mode == _lib.LOSS_FIXMATCH:
            conf_r, lab_r, mask_r = O.pseudo_label(zt, thr)
            loss_u = O.ce_masked(zs[Bl:], lab_r, mask_r)
            mratio = float(mask_r.float().mean())
        elif mode == _lib.LOSS_SOFT_MASKED:     # the consistency term of ReCo (reco.py:226, 248-250)
            conf_r, lab_r, mask_r = O.pseudo_label(zt, thr)
            loss_u = O.ce_soft_masked(zs[Bl:], zt.double().softmax(1), mask_r)
            mratio = float(mask_r.float().mean())
        else:
            loss_u = O.ce_soft(zs[Bl:], zt.double().softmax(1))
        loss = (loss_x + loss_u) / 2
    loss.backward()
    dlow = torch.full((S, Lin, ncls), 9.0, device=DEV)
    sums = torch.zeros(4, dtype=torch.float64, device=DEV)
    conf = torch.zeros(max(Bu, 1), L, device=DEV)
    label = torch.zeros(max(Bu, 1), L, dtype=torch.int64, device=DEV)
    mask = torch.zeros(max(Bu, 1), L, dtype=torch.uint8, device=DEV)
    mat = mode in (_lib.LOSS_FIXMATCH, _lib.LOSS_SOFT_MASKED)
    call("ssb_semi_loss", low_s.data_ptr(), y.data_ptr(), low_t.data_ptr() if Bu else None, dlow.data_ptr(), sums.data_ptr(),
         Bl, Bu, Lin, L, ncls, mode, thr, None, 0, conf.data_ptr() if mat else None, label.data_ptr() if mat else None,
         mask.data_ptr() if mat else None, st())
    s = sums.cpu()
    assert abs(float(s[0]) / (Bl * L) - float(loss_x)) < 2e-6 * max(1, abs(float(loss_x)))
    if mode != _lib.LOSS_SUP:
        assert abs(float(s[1]) / (Bu * L) - float(loss_u)) < 5e-6 * max(1, abs(float(loss_u)))
    if mode == _lib.LOSS_FIXMATCH:
        # masks computed from kernel-internal fp32 lerp may differ from the torch path at exact threshold ties only
        assert abs(float(s[2]) / (Bu * L) - mratio) < 1e-4
        assert float((label != lab_r).float().mean()) < 1e-4
        assert float((mask.bool() != mask_r).float().mean()) < 1e-4
        assert float((conf - conf_r).abs().max()) < 5e-5
    if mode == _lib.LOSS_SOFT_MASKED:
        assert 0.1 < mratio < 0.9 and abs(float(s[2]) / (Bu * L) - mratio) < 1e-4
        assert float((mask.bool() != mask_r).float().mean()) < 1e-4 and float((conf - conf_r).abs().max()) < 5e-5
    assert rel_err(dlow.permute(0, 2, 1), ls.grad) < 2e-5


def test_semi_loss_deterministic():
    torch.manual_seed(0)
    Bl, Bu, Lin, L, ncls = 4, 4, 79, 2500, 4
    low_s = torch.randn(Bl + Bu, Lin, ncls, device=DEV)
    low_t = torch.randn(Bu, Lin, ncls, device=DEV) * 4
    y = torch.randint(0, ncls, (Bl, L), device=DEV)
    outs = []
    for _ in range(3):
        dlow = torch.zeros(Bl + Bu, Lin, ncls, device=DEV)
        sums = torch.zeros(4, dtype=torch.float64, device=DEV)
        call("ssb_semi_loss", low_s.data_ptr(), y.data_ptr(), low_t.data_ptr(), dlow.data_ptr(), sums.data_ptr(), Bl, Bu, Lin, L,
             ncls, _lib.LOSS_FIXMATCH, 0.7, None, 0, None, None, None, st())
        outs.append(dlow)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_adamw_ema_matches_oracle():
    torch.manual_seed(0)
    n = 64 * 1024 + 64
    p = torch.randn(n, device=DEV)
    m = torch.randn(n, device=DEV) * 0.01
    v = torch.rand(n, device=DEV) * 1e-4
    e = torch.randn(n, device=DEV)
    g = torch.randn(n, device=DEV) * 0.1
    pr, mr, vr, er = p.double().clone(), m.double().clone(), v.double().clone(), e.double().clone()
    lr, b1, b2, eps, wd, t, d = 3e-4, 0.9, 0.999, 1e-8, 0.05, 7, 0.99
    O.adamw_update(pr, g.double(), mr, vr, t, lr, (b1, b2), eps, wd)
    er = er * d + pr * (1 - d)
    sp = StepParams()
    sp.lr, sp.inv_bias1, sp.inv_sqrt_bias2 = lr, 1 / (1 - b1 ** t), 1 / (1 - b2 ** t) ** 0.5
    sp.ema_decay, sp.ema_first, sp.step, sp.grad_scale = d, 0, t, 1.0
    spd = torch.frombuffer(bytearray(bytes(sp)), dtype=torch.uint8).to(DEV)
    call("ssb_adamw_ema", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), e.data_ptr(), n, b1, b2, eps, wd, spd.data_ptr(), st())
    assert rel_err(p, pr) < 1e-6 and rel_err(m, mr) < 1e-6 and rel_err(v, vr) < 1e-6 and rel_err(e, er) < 1e-6
    # teacher-aliasing first step: EMA result == new student weights (up to one rounding)
    sp.ema_first = 1
    spd = torch.frombuffer(bytearray(bytes(sp)), dtype=torch.uint8).to(DEV)
    call("ssb_adamw_ema", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), e.data_ptr(), n, b1, b2, eps, wd, spd.data_ptr(), st())
    assert rel_err(e, p) < 1e-6
    # grad norm
    ws = torch.zeros(1, dtype=torch.float64, device=DEV)
    out = torch.zeros(1, device=DEV)
    call("ssb_grad_norm", g.data_ptr(), n, ws.data_ptr(), out.data_ptr(), st())
    assert abs(float(out) - float(g.double().norm())) < 1e-4 * float(g.double().norm())


def test_syncbn_exchange_single_gpu():
    """ssb_syncbn_exchange with two mailboxes on ONE GPU: the peer's words are pre-filled (nothing waits on another
    kernel), so the launch must (a) deposit this rank's slice in the peer's mailbox in the wire format and (b) replace
    its slice by the rank-ordered sum."""
    lib = _lib.load()
    n, slot, world = 200, 256, 2
    nbytes = int(lib.ssb_syncbn_mailbox_bytes(slot))
    box = [torch.zeros(nbytes, dtype=torch.uint8, device=DEV) for _ in range(world)]
    peers = torch.tensor([b.data_ptr() for b in box], dtype=torch.int64, device=DEV)
    torch.manual_seed(3)
    mine = torch.randn(n, dtype=torch.float64, device=DEV)
    theirs = torch.randn(n, dtype=torch.float64, device=DEV)
    epoch = 1
    # wire format: per double two 8-byte words {lo32, epoch}, {hi32, epoch}; parity = epoch & 1; slot of the sender rank
    words = torch.zeros(n, 4, dtype=torch.int32, device=DEV)
    raw = theirs.view(torch.int32).view(n, 2)
    words[:, 0], words[:, 1], words[:, 2], words[:, 3] = raw[:, 0], epoch, raw[:, 1], epoch
    off = 4096 + ((epoch & 1) * 16 + 1) * slot * 16          # rank 0's mailbox, slot of rank 1
    box[0][off: off + n * 16] = words.view(torch.uint8).flatten()
    sl = mine.clone()
    call("ssb_syncbn_exchange", sl.data_ptr(), n, peers.data_ptr(), world, 0, slot, st())
    torch.cuda.synchronize()
    assert torch.equal(sl, mine + theirs)
    off1 = 4096 + ((epoch & 1) * 16 + 0) * slot * 16         # rank 1's mailbox, slot of rank 0
    got = box[1][off1: off1 + n * 16].view(torch.int32).view(n, 4)
    assert torch.equal(got[:, [0, 2]].contiguous().view(torch.float64).flatten(), mine) and bool((got[:, [1, 3]] == epoch).all())
    assert int(box[0][:4].view(torch.int32)[0]) == 1 and int(box[0][64:68].view(torch.int32)[0]) == 0   # counter, no timeout


def _fused_sync_setup(bn, t, Cn, world, step, fwd_vals=None, bwd_vals=None, n_sums=None):
    """Two mailboxes on ONE GPU for the in-kernel SyncBN exchange (ssb_bn.sync_*): this process is rank 0, the values
    rank 1 'sent' are pre-filled in rank 0's mailbox in the wire format, tagged with `step`.  Returns what must stay alive."""
    lib = _lib.load()
    n_sums = n_sums or 2 * Cn + 64
    slot = 2 * n_sums
    nbytes = int(lib.ssb_syncbn_fused_mailbox_bytes(slot, world))
    box = [torch.zeros(nbytes, dtype=torch.uint8, device=DEV) for _ in range(world)]
    peers = torch.tensor([b.data_ptr() for b in box], dtype=torch.int64, device=DEV)
    sp = StepParams()
    sp.step = step
    sp_dev = torch.frombuffer(bytearray(bytes(sp)), dtype=torch.uint8).to(DEV)
    bn.sync_peers, bn.sync_sp = peers.data_ptr(), sp_dev.data_ptr()
    bn.sync_world, bn.sync_rank, bn.sync_slot = world, 0, slot
    bn.sync_fwd_off, bn.sync_bwd_off = 8, n_sums + 8
    for vals, off in ((fwd_vals, 8), (bwd_vals, n_sums + 8)):
        if vals is None:
            continue
        n = vals.numel()
        words = torch.zeros(n, 4, dtype=torch.int32, device=DEV)
        raw = vals.contiguous().view(torch.int32).view(n, 2)
        words[:, 0], words[:, 1], words[:, 2], words[:, 3] = raw[:, 0], step, raw[:, 1], step
        o = 4096 + (1 * slot + off) * 16                     # rank 0's mailbox, region of rank 1
        box[0][o: o + n * 16] = words.view(torch.uint8).flatten()
    return box, peers, sp_dev, slot


def _pushed(box, slot, off, n, step):
    """what rank 0 deposited in rank 1's mailbox (region of rank 0): values, and whether every word carries the tag"""
    o = 4096 + (0 * slot + off) * 16
    got = box[1][o: o + n * 16].view(torch.int32).view(n, 4)
    return got[:, [0, 2]].contiguous().view(torch.float64).flatten(), bool((got[:, [1, 3]] == step).all())


@pytest.mark.parametrize("dtype", [_lib.F32, _lib.BF16])
def test_syncbn_in_kernel_exchange_forward(dtype):
    """ssb_bn_act_fwd (train) with ssb_bn.sync_* set: this rank's statistics go to the peer's mailbox, the peer's
    (pre-filled) are added in rank order, and the output / saved mean, invstd / running statistics are those of the
    layer run on the TOTAL statistics with count_mul = 2."""
    torch.manual_seed(11)
    B, Cn, L, pitch = 6, 128, 157, 160
    g = Geom(B, pitch, L, Cn)
    xx = to_flat(torch.randn(B, Cn, L, device=DEV, dtype=torch.float64) * 1.3 + 0.2, pitch, dtype)
    theirs = torch.cat((torch.randn(Cn, dtype=torch.float64, device=DEV) * 30, torch.rand(Cn, dtype=torch.float64, device=DEV) * 2000 + 1500))
    outs = []
    for sync in (False, True):
        torch.manual_seed(98)            # same gamma / beta / running statistics in both runs
        bn, t = make_bn(Cn, train_count_mul=2)
        call("ssb_bn_stats", xx.data_ptr(), g, t["sums"].data_ptr(), dtype, st())
        local = t["sums"].clone()
        keep = None
        if sync:
            keep = _fused_sync_setup(bn, t, Cn, 2, 7, fwd_vals=theirs)
        else:
            t["sums"] += theirs
        y = torch.full((B * pitch, Cn), 5.0, dtype=TDT[dtype], device=DEV)
        call("ssb_bn_act_fwd", xx.data_ptr(), C.byref(bn), None, None, y.data_ptr(), g, 1, 1, dtype, st())
        torch.cuda.synchronize()
        outs.append((y.clone(), t["mi"].clone(), t["rm"].clone(), t["rv"].clone()))
        if sync:
            box, _, _, slot = keep
            vals, tagged = _pushed(box, slot, 8, 2 * Cn, 7)
            assert tagged and torch.equal(vals, local)
            assert int(box[0][64:68].view(torch.int32)[0]) == 0      # no timeout recorded
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("dtype", [_lib.F32, _lib.BF16])
@pytest.mark.parametrize("fused", [True, False])
def test_syncbn_in_kernel_exchange_backward(dtype, fused):
    """ssb_bn_bwd_fused / ssb_bn_bwd_apply with ssb_bn.sync_* set == the two-launch path with the peer's sums added by hand."""
    torch.manual_seed(12)
    B, Cn, L, pitch = 16, 128, 313, 316
    g = Geom(B, pitch, L, Cn)
    gr = to_flat(torch.randn(B, Cn, L, device=DEV, dtype=torch.float64), pitch, dtype)
    yy = to_flat(torch.relu(torch.randn(B, Cn, L, device=DEV, dtype=torch.float64)), pitch, dtype)
    xx = to_flat(torch.randn(B, Cn, L, device=DEV, dtype=torch.float64) + 0.3, pitch, dtype)
    theirs = torch.randn(2 * Cn, dtype=torch.float64, device=DEV) * 40
    res = []
    for sync in (False, True):
        torch.manual_seed(99)
        bn, t = make_bn(Cn, train_count_mul=2)
        t["mi"][:Cn] = 0.1
        t["mi"][Cn:] = 0.9
        dx = torch.full((B * pitch, Cn), 3.0, dtype=TDT[dtype], device=DEV)
        keep = _fused_sync_setup(bn, t, Cn, 2, 3, bwd_vals=theirs) if sync else None
        if sync and fused:
            bar = torch.zeros(1, dtype=torch.int32, device=DEV)
            rep = torch.zeros(8, 2 * Cn + 64, dtype=torch.float64, device=DEV)
            call("ssb_bn_bwd_fused", gr.data_ptr(), yy.data_ptr(), xx.data_ptr(), C.byref(bn), dx.data_ptr(), None, None, None, None, g,
                 bar.data_ptr(), rep.data_ptr(), None, 2 * Cn + 64, dtype, st())
        else:
            call("ssb_bn_bwd_reduce", gr.data_ptr(), None, yy.data_ptr(), xx.data_ptr(), C.byref(bn), None, None, g, dtype, st())
            if not sync:
                t["bsums"] += theirs
            call("ssb_bn_bwd_apply", gr.data_ptr(), None, yy.data_ptr(), xx.data_ptr(), C.byref(bn), dx.data_ptr(), None, None,
                 None, None, g, dtype, st())
        torch.cuda.synchronize()
        res.append((dx.float(), t["dgamma"].clone(), t["dbeta"].clone()))
        if sync:
            box, _, _, slot = keep
            vals, tagged = _pushed(box, slot, (2 * Cn + 64) + 8, 2 * Cn, 3)
            assert tagged and int(box[0][64:68].view(torch.int32)[0]) == 0
    tol = 1e-5 if dtype == _lib.F32 else 1e-2
    assert rel_err(res[1][0], res[0][0]) < tol and halo_is_zero(res[1][0], B, pitch, L)
    assert rel_err(res[1][1], res[0][1]) < 1e-5 and rel_err(res[1][2], res[0][2]) < 1e-5


def test_errors_are_reported():
    lib = _lib.load()
    g = Geom(1, 4, 8, 8)   # pitch < len + 2
    rc = lib.ssb_bn_stats(None, g, None, 0, None)
    assert rc != 0 and b"ssb_bn_stats" in lib.ssb_last_error()
    with pytest.raises(RuntimeError):
        call("ssb_conv1d_fwd", None, None, None, Geom(1, 10, 8, 8), Geom(1, 10, 8, 8), 5, 1, 0, 0, None)
