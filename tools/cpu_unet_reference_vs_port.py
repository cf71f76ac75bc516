import sys, time, json, os, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from oracle import unet_oracle as uo
sys.path.insert(0, "/root/reference")
from holo_diffusion.guided_diffusion.unet import UNetModel
torch.set_num_threads(os.cpu_count())
C, R = 32, 32
sd = uo.make_unet_state_dict(C, C, seed=2)
ref = UNetModel(dims=3, image_size=R, in_channels=C, model_channels=64, out_channels=C, num_res_blocks=2,
                attention_resolutions=[4, 8], dropout=0.0, channel_mult=[1, 1, 2, 4, 8], num_classes=None, use_checkpoint=False,
                num_heads=2, num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=True, resblock_updown=False,
                zero_last_conv=False, homogeneous_resample=True)
ref.load_state_dict(sd, strict=True); ref.eval()
x = torch.tanh(torch.randn(1, C, R, R, R, generator=torch.Generator().manual_seed(0)))
t = torch.zeros(1, dtype=torch.long)
def timeit(f, n=3):
    f(); ts = []
    for _ in range(n):
        s = time.perf_counter(); y = f(); ts.append(time.perf_counter() - s)
    return sorted(ts)[len(ts)//2], y
with torch.no_grad():
    tr, yr = timeit(lambda: ref(x, timesteps=t))
    tp, yp = timeit(lambda: uo.unet_forward(sd, x, t))
print(json.dumps({"what": "base-args UNet forward at 32^3 x 32ch on this container's host cores: the reference's own UNetModel (imported from /root/reference) vs the oracle port bench.py times", "threads": os.cpu_count(), "reference_UNetModel_s": tr, "oracle_port_s": tp, "ratio_port_over_reference": tp / tr, "outputs_bit_identical": bool(torch.equal(yr, yp))}))
