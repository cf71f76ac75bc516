"""Where does a small tcgen05 convolution launch spend its time?  (run under gpurun)
1. clock64 stamps of CTA 0 (holo_debug_conv_trace) for coarse-level shapes of the base UNet;
2. a chain of 24 dependent coarse-level convolutions timed as eager launches and as one CUDA graph: the per-launch cost
   inside the replayed step (HOLO_PDL=1 in the environment switches programmatic dependent launch on)."""
import ctypes
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from holo_diffusion_b200 import ops
from holo_diffusion_b200._lib import lib

NAMES = ["entry", "setup done", "first TMA issued", "first operands landed", "last MMA issued", "first accumulator ready",
         "first item written", "exit"]


def operands(Cin, Cout, dims, k):
    V = dims[0] * dims[1] * dims[2]
    g = torch.Generator(device="cuda").manual_seed(0)
    hi = torch.randn(V, Cin, device="cuda", generator=g).to(torch.float16)
    lo = (torch.randn(V, Cin, device="cuda", generator=g) * 0.0004).to(torch.float16)
    w_hi = (torch.randn(Cout, k ** 3, Cin, device="cuda", generator=g) / math.sqrt(Cin * k ** 3) * 512).to(torch.float16)
    w_lo = (w_hi.float() * 0.0004).to(torch.float16)
    return hi, lo, w_hi, w_lo, torch.zeros(Cout, device="cuda"), torch.empty(V, Cout, device="cuda")


def trace(Cin, Cout, dims, k=3):
    hi, lo, w_hi, w_lo, b, out = operands(Cin, Cout, dims, k)
    buf = torch.zeros(8, dtype=torch.int64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        ops.conv3d_tc(hi, lo, Cin, dims, k, w_hi, w_lo, b, None, Cout, out, w_scale=512.0)
    lib().cdll.holo_debug_conv_trace(ctypes.c_void_p(buf.data_ptr()))
    rows = []
    for _ in range(5):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ops.conv3d_tc(hi, lo, Cin, dims, k, w_hi, w_lo, b, None, Cout, out, w_scale=512.0)
        e.record()
        torch.cuda.synchronize()
        t = buf.cpu().tolist()
        rows.append(([(x - t[0]) for x in t], s.elapsed_time(e) * 1e3))
    lib().cdll.holo_debug_conv_trace(None)
    rows.sort(key=lambda r: r[1])
    t, us = rows[len(rows) // 2]
    print(f"Cin={Cin} Cout={Cout} dims={dims} k={k}: event {us:.1f} us; CTA 0 cycles since entry (us at 1.9 GHz): "
          + ", ".join(f"{n} {c} ({c / 1900:.2f})" for n, c in zip(NAMES[1:], t[1:])), flush=True)


def chain(n=24):
    shapes = [(256, 256, (8, 8, 8)), (512, 512, (4, 4, 4)), (128, 128, (16, 16, 16))]
    res = {}
    for Cin, Cout, dims in shapes:
        hi, lo, w_hi, w_lo, b, out = operands(Cin, Cout, dims, 3)
        V = out.shape[0]
        o_hi = torch.empty(V, Cout, device="cuda", dtype=torch.float16)
        o_lo = torch.empty_like(o_hi)

        def step():   # conv -> split (stands for the GroupNorm pass) -> conv ...: n dependent pairs
            x_hi, x_lo = hi, lo
            for _ in range(n):
                ops.conv3d_tc(x_hi, x_lo, Cin, dims, 3, w_hi, w_lo, b, None, Cout, out, w_scale=512.0)
                ops.split_bf16(out, V, Cout, Cout, o_hi, o_lo)
                x_hi, x_lo = o_hi, o_lo

        for _ in range(3):
            step()
        torch.cuda.synchronize()

        def timeit(f, reps=10):
            ts = []
            for _ in range(reps):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda._sleep(20_000_000)   # device-bound: the host queues the chain behind a spin
                s.record(); f(); e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e) * 1e3)
            ts.sort()
            return ts[len(ts) // 2]

        eager = timeit(step)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        graph = timeit(g.replay)
        res[f"{Cin}->{Cout}@{dims[0]}^3"] = {"eager_us_per_pair": eager / n, "graph_us_per_pair": graph / n}
        print(f"chain {Cin}->{Cout}@{dims}: eager {eager / n:.2f} us per (conv + split), graph {graph / n:.2f}", flush=True)
    return res


if __name__ == "__main__":
    print("HOLO_PDL =", os.environ.get("HOLO_PDL"), " HOLO_CONV_CHUNK =", os.environ.get("HOLO_CONV_CHUNK"))
    if "--no-trace" not in sys.argv:
        trace(64, 64, (4, 4, 4), 1)
        trace(512, 512, (4, 4, 4))
        trace(256, 256, (8, 8, 8))
        trace(128, 128, (16, 16, 16))
        trace(64, 64, (32, 32, 32))
    chain()
