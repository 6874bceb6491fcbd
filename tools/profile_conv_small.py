"""One launch of each direct conv kernel (csrc/conv_small.cu) at FashionMNIST B=4096 shapes, for ncu captures.
   python tools/profile_conv_small.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from multimodal_vae_public_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = "cuda"
x = torch.rand(B, 784, device=dev); wc = torch.randn(64, 16, device=dev) * 0.25
a = torch.empty(B * 196, 64, device=dev); h = torch.empty_like(a)
da = torch.randn(B * 196, 64, device=dev); dwc = torch.zeros(64, 16, device=dev)
B2 = 2 * B
hin = torch.randn(B2 * 196, 64, device=dev); ain = torch.randn(B2 * 196, 64, device=dev); wt = torch.randn(16, 64, device=dev) * 0.1
out = torch.empty(B2, 28, 28, 1, device=dev); dout = torch.randn(B2, 28, 28, 1, device=dev)
dhin = torch.empty_like(hin); dwt = torch.zeros(16, 64, device=dev)
for it in range(3):
    ops.conv_cin_fwd(x, wc, a, h, B, 28, 28, 1, 64)
    ops.conv_cin_wgrad(x, da, dwc, B, 28, 28, 1, 64)
    ops.convT_cout_fwd(hin, wt, out, B2, 14, 14, 64, 1)
    ops.convT_cout_bwd(dout, hin, ain, wt, dhin, dwt, B2, 14, 14, 64, 1)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
ev[0].record(); ops.conv_cin_fwd(x, wc, a, h, B, 28, 28, 1, 64)
ev[1].record(); ops.conv_cin_wgrad(x, da, dwc, B, 28, 28, 1, 64)
ev[2].record(); ops.convT_cout_fwd(hin, wt, out, B2, 14, 14, 64, 1)
ev[3].record(); ops.convT_cout_bwd(dout, hin, ain, wt, dhin, dwt, B2, 14, 14, 64, 1)
ev[4].record(); torch.cuda.synchronize()
names = ["conv_cin_fwd", "conv_cin_wgrad", "convT_cout_fwd", "convT_cout_bwd"]
bytes_ = [2 * a.numel() * 4 + x.numel() * 4, da.numel() * 4 + x.numel() * 4, hin.numel() * 4 + out.numel() * 4,
          3 * hin.numel() * 4 + dout.numel() * 4]
for i, n in enumerate(names):
    ms = ev[i].elapsed_time(ev[i + 1])
    print(f"{n:16s} {ms * 1e3:8.1f} us   {bytes_[i] / ms / 1e6:7.1f} GB/s (algorithmic bytes {bytes_[i] / 1e6:.0f} MB)")
