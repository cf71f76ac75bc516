"""Micro-benchmark of single convolution layers on the tcgen05 kernels (run under gpurun)."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from holo_diffusion_b200 import ops

def run(Cin, Cout, dims, k=3, iters=20):
    D, H, W = dims
    V = D * H * W
    g = torch.Generator(device="cuda").manual_seed(0)
    hi = torch.randn(V, Cin, device="cuda", generator=g).to(torch.bfloat16)
    lo = (torch.randn(V, Cin, device="cuda", generator=g) * 0.004).to(torch.bfloat16)
    w_hi = (torch.randn(Cout, k ** 3, Cin, device="cuda", generator=g) / math.sqrt(Cin * k ** 3)).to(torch.bfloat16)
    w_lo = (w_hi.float() * 0.004).to(torch.bfloat16)
    b = torch.zeros(Cout, device="cuda")
    out = torch.empty(V, Cout, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ops.conv3d_tc(hi, lo, Cin, dims, k, w_hi, w_lo, b, None, Cout, out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops.conv3d_tc(hi, lo, Cin, dims, k, w_hi, w_lo, b, None, Cout, out); e.record()
        torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort()
    ms = ts[len(ts) // 2]
    fl = 2.0 * V * Cout * Cin * k ** 3
    print(f"Cin={Cin:4d} Cout={Cout:4d} dims={dims} k={k}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TF/s alg  {3*fl/ms/1e9:7.1f} TF/s exec", flush=True)

if __name__ == "__main__":
    print("HOLO_CONV_NO_HALO =", os.environ.get("HOLO_CONV_NO_HALO"))
    run(64, 64, (64, 64, 64))
    run(128, 64, (64, 64, 64))
    run(64, 64, (32, 32, 32))
    run(128, 128, (16, 16, 16))
    run(256, 256, (8, 8, 8))
    run(128, 64, (64, 64, 64), k=1)
