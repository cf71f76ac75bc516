"""Where does the denoiser's distance to the reference come from?  Runs the base-args UNet on a synthetic grid
through our kernels and through the oracle restatement (torch ops on the same GPU, fp32 with TF32 off, or fp64 with
--f64) and prints the relative error after every block (input_blocks.i / middle_block / output_blocks.i / out).
Test-side diagnostic (it executes oracle/, so it lives under tests/): never part of the product path or of bench.py.

  python tests/diagnostics/unet_error_trace.py --resol 64 --channels 32 [--f64] [--no-tc]
  HOLO_PAIR_FMT=bf16 python tests/diagnostics/unet_error_trace.py ...      # bf16 operand pairs
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import unet_oracle as uo  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--resol", type=int, default=64)
    ap.add_argument("--channels", type=int, default=32)
    ap.add_argument("--f64", action="store_true")
    ap.add_argument("--no-tc", action="store_true")
    a = ap.parse_args()
    from holo_diffusion_b200.unet import SimpleUnet3D, UNetExecutor
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    C, R = a.channels, a.resol
    kw = dict(model_channels=64, num_res_blocks=2, channel_mult=(1, 1, 2, 4, 8), attention_resolutions=(4, 8), num_heads=2)
    sd = uo.make_unet_state_dict(C, C, seed=2)
    net = SimpleUnet3D(image_size=R, in_channels=C, out_channels=C, use_tensor_cores=not a.no_tc, **kw)
    net._net.load_state_dict(sd, strict=True)
    net.cuda()
    x = torch.tanh(torch.randn(1, C, R, R, R, generator=torch.Generator().manual_seed(0))).cuda()
    t = torch.zeros(1, dtype=torch.long, device="cuda")
    ours, ref = [], []
    orig_run = UNetExecutor._run

    def run(self, seq, act, film_all):
        r = orig_run(self, seq, act, film_all)
        D, H, W = r.dims
        ours.append(r.x1.t().reshape(1, r.C, D, H, W).clone())
        return r

    UNetExecutor._run = run
    y = net(x, t)
    UNetExecutor._run = orig_run
    orig_block = uo._run_block

    def block(sd_, pre, h, emb, n_heads):
        r = orig_block(sd_, pre, h, emb, n_heads)
        ref.append((pre, r))
        return r

    uo._run_block = block
    with torch.no_grad():
        if a.f64:
            yr = uo.unet_forward({k: v.cuda().double() for k, v in sd.items()}, x.double(), t)
        else:
            yr = uo.unet_forward({k: v.cuda() for k, v in sd.items()}, x, t)
    uo._run_block = orig_block
    torch.cuda.synchronize()

    def rel(p, q):
        return float((p.double() - q.double()).abs().max() / q.double().abs().max())

    assert len(ours) == len(ref), (len(ours), len(ref))
    print(f"pairs={net._exec.pair_dtype} tc={not a.no_tc} ref={'fp64' if a.f64 else 'fp32'} grid {R}^3 x {C}")
    for o, (pre, r) in zip(ours, ref):
        print(f"{pre:20s} {tuple(r.shape[1:])!s:24s} rel err {rel(o, r):.2e}   max|ref| {float(r.abs().max()):.3g}")
    print(f"{'out':20s} {tuple(yr.shape[1:])!s:24s} rel err {rel(y, yr):.2e}   max|ref| {float(yr.abs().max()):.3g}")
    print(f"tanh(out): rel err {rel(torch.tanh(y), torch.tanh(yr)):.2e}")


if __name__ == "__main__":
    main()
