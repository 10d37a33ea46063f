"""Times ssb_stem_conv_wgrad alone (config-2 shape: 32 strips x 1 lead x 2500 samples, 64 stem channels, bf16)."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "semi-seg-ecg_b200", "src"))
import torch  # noqa: E402

from semiseg_b200 import _lib  # noqa: E402
from semiseg_b200._lib import Geom, call  # noqa: E402

B, Cl, L, Cs = 32, int(os.environ.get("LEADS", "1")), 2500, 64
L0 = 1250
p0 = 2 * (2 * (313 + 2))          # any pitch >= L0 + 2 that is a multiple of 4
g = Geom(B, p0, L0, Cs)
x = torch.randn(B, Cl, L, device="cuda")
dy = torch.randn(B * p0, Cs, device="cuda").bfloat16()
dw = torch.zeros(Cs, Cl, 7, device="cuda")
_lib.check(_lib.load().ssb_prepare(), "prepare") if hasattr(_lib.load(), "ssb_prepare") else None
st = torch.cuda.current_stream().cuda_stream
for _ in range(20):
    call("ssb_stem_conv_wgrad", x.data_ptr(), dy.data_ptr(), dw.data_ptr(), Cl, L, g, _lib.BF16, st)
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    for _ in range(50):
        call("ssb_stem_conv_wgrad", x.data_ptr(), dy.data_ptr(), dw.data_ptr(), Cl, L, g, _lib.BF16, torch.cuda.current_stream().cuda_stream)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
gr.replay()
torch.cuda.synchronize()
a.record()
for _ in range(10):
    gr.replay()
b.record()
torch.cuda.synchronize()
print(f"SSB_STEM_WGRAD_DIRECT={os.environ.get('SSB_STEM_WGRAD_DIRECT', '1')} leads={Cl}: {a.elapsed_time(b) / 500 * 1e3:.2f} us per launch (back to back in a graph)")
