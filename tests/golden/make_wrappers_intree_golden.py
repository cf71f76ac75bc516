"""Golden facts from the reference's IN-TREE wrapper code, executed UNMODIFIED from /root/reference against the
pytorch3d stand-in (oracle/pt3d_stub; see its README for what that does and does not pin):

  * ``SimpleUnet3D`` (utils/diffusion_utils.py:41-86): state-dict keys / shapes, initialisation facts (which biases
    are zero, which tensors Xavier re-initialises, zeroed proj_out), ``forward`` with ``cond_features``.
  * ``ImplicitronGaussianDiffusion`` (utils/diffusion_utils.py:89-140): the schedule tables its DEFAULTS produce.
  * ``get_simple_360_camera_trajectory`` (utils/render_utils/flyaround.py:301-350): its source is read from the
    reference file and executed as is (the module itself drags in the dataset / video stack), which pins the azimuth
    / elevation conversion and the ORDER of the up-axis correction R = R_plane @ R_lookat.

    python tests/golden/make_wrappers_intree_golden.py     # writes tests/golden/wrappers_intree_ref.npz (+ .json)
"""
import ast
import json
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
OUT = os.path.join(HERE, "wrappers_intree_ref.npz")
OUT_JSON = os.path.join(HERE, "wrappers_intree_ref.json")
UNET = dict(image_size=16, in_channels=16, out_channels=16, model_channels=64, num_res_blocks=2,
            channel_mult=(1, 1, 2, 4, 8), attention_resolutions=(4, 8), num_heads=2)


def generate():
    for p in (ROOT, REF, os.path.join(ROOT, "oracle", "pt3d_stub")):   # REF ahead of ROOT: holo_diffusion = the reference, not the shim
        if p not in sys.path:
            sys.path.insert(0, p)
    from holo_diffusion.utils.diffusion_utils import ImplicitronGaussianDiffusion, SimpleUnet3D

    arrays, facts = {}, {}
    # ---- SimpleUnet3D
    torch.manual_seed(0)
    net = SimpleUnet3D(**UNET)
    sd = net.state_dict()
    facts["unet_keys"] = {k: list(v.shape) for k, v in sd.items()}
    zero_bias, nonzero_bias, zero_weight, within_xavier = [], [], [], []
    for name, m in net._net.named_modules():
        if isinstance(m, (torch.nn.Conv3d, torch.nn.Linear, torch.nn.Conv1d)):
            (zero_bias if float(m.bias.detach().abs().max()) == 0.0 else nonzero_bias).append(name)
            if float(m.weight.detach().abs().max()) == 0.0:
                zero_weight.append(name)
            elif isinstance(m, (torch.nn.Conv3d, torch.nn.Linear)):
                rf = int(np.prod(m.weight.shape[2:])) if m.weight.dim() > 2 else 1
                bound = math.sqrt(6.0 / ((m.weight.shape[0] + m.weight.shape[1]) * rf))
                if float(m.weight.detach().abs().max()) <= bound * (1 + 1e-6):
                    within_xavier.append(name)
    facts["unet_init"] = {"zero_bias": zero_bias, "nonzero_bias": nonzero_bias, "zero_weight": zero_weight,
                          "within_xavier_bound": within_xavier}
    # forward with cond_features: x and cond are concatenated on the channel axis (in_channels = 8 + 8)
    # (a small architecture keeps the vectors small: 8^3 grid, 2 levels)
    from oracle import unet_oracle as uo
    small = SimpleUnet3D(image_size=8, in_channels=8, out_channels=8, model_channels=32, num_res_blocks=1,
                         channel_mult=(1, 2), attention_resolutions=(2,), num_heads=1)
    fix = uo.make_unet_state_dict(8, 8, 32, 1, (1, 2), (2,), seed=5)
    small._net.load_state_dict(fix, strict=True)
    g = torch.Generator().manual_seed(4)
    x, c = torch.randn(1, 5, 8, 8, 8, generator=g), torch.randn(1, 3, 8, 8, 8, generator=g)
    with torch.no_grad():
        y = small(x, torch.full((1,), 37, dtype=torch.long), cond_features=c)
    arrays.update({"unet/x": x.numpy(), "unet/cond": c.numpy(), "unet/y": y.numpy()})
    # ---- ImplicitronGaussianDiffusion defaults
    d = ImplicitronGaussianDiffusion()._diffusion
    for k in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
              "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
              "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2"):
        arrays["diffusion/" + k] = np.asarray(getattr(d, k), dtype=np.float64)
    facts["diffusion"] = {"num_timesteps": int(d.num_timesteps), "model_mean_type": d.model_mean_type.name,
                          "model_var_type": d.model_var_type.name, "rescale_timesteps": bool(d.rescale_timesteps)}
    # ---- get_simple_360_camera_trajectory: the function's own source, executed in a namespace with the stand-ins
    src = open(os.path.join(REF, "holo_diffusion/utils/render_utils/flyaround.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "get_simple_360_camera_trajectory")
    from typing import Tuple

    from pytorch3d.renderer import PerspectiveCameras, look_at_view_transform
    ns = {"torch": torch, "np": np, "math": math, "Tuple": Tuple, "look_at_view_transform": look_at_view_transform,
          "PerspectiveCameras": PerspectiveCameras}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "flyaround.py", "exec"), ns)
    up = (-0.0396, -0.8306, -0.5554)   # visualize_reconstruction.py:35
    for n_poses, max_angle in ((8, 2 * math.pi), (5, math.pi)):
        cams = ns["get_simple_360_camera_trajectory"](max_angle, n_poses, -math.pi / 6, 10.0, up, 3.2)
        tag = f"cams{n_poses}"
        arrays[tag + "/R"], arrays[tag + "/T"] = cams.R.numpy(), cams.T.numpy()
        arrays[tag + "/focal"], arrays[tag + "/pp"] = cams.focal_length.numpy(), cams.principal_point.numpy()
    # ---- call signatures of the plug-in surface, read from the reference SOURCE
    facts["signatures"] = {}
    for rel, cls, fn_name in (("holo_diffusion/holo_diffusion_model.py", "HoloDiffusionModel", "forward"),
                              ("holo_diffusion/holo_voxel_grid_implicit_function.py", "HoloVoxelGridImplicitFunction", "forward"),
                              ("holo_diffusion/holo_voxel_grid_implicit_function.py", "RenderMLP", "forward"),
                              ("holo_diffusion/holo_multipass_ea.py", "HoloMultiPassEmissionAbsorptionRenderer", "_run_raymarcher"),
                              ("holo_diffusion/utils/diffusion_utils.py", "Unet3DBase", "forward"),
                              ("holo_diffusion/utils/diffusion_utils.py", "SimpleUnet3D", "forward"),
                              ("holo_diffusion/utils/render_utils/flyaround.py", None, "get_simple_360_camera_trajectory")):
        tree = ast.parse(open(os.path.join(REF, rel)).read())
        body = tree.body if cls is None else next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
        f = next(n for n in body if isinstance(n, ast.FunctionDef) and n.name == fn_name)
        facts["signatures"][(cls + "." if cls else "") + fn_name] = signature_of(f.args)
    return arrays, facts


def signature_of(a: ast.arguments):
    """[(name, kind, default-source-or-None)] with kinds as in inspect.Parameter."""
    out = []
    pos = a.posonlyargs + a.args
    d0 = len(pos) - len(a.defaults)
    for i, arg in enumerate(pos):
        out.append([arg.arg, "POSITIONAL_OR_KEYWORD", ast.unparse(a.defaults[i - d0]) if i >= d0 else None])
    if a.vararg:
        out.append([a.vararg.arg, "VAR_POSITIONAL", None])
    for arg, d in zip(a.kwonlyargs, a.kw_defaults):
        out.append([arg.arg, "KEYWORD_ONLY", ast.unparse(d) if d is not None else None])
    if a.kwarg:
        out.append([a.kwarg.arg, "VAR_KEYWORD", None])
    return out


if __name__ == "__main__":
    a, f = generate()
    np.savez_compressed(OUT, **a)
    json.dump(f, open(OUT_JSON, "w"), indent=0, sort_keys=True)
    print(f"wrote {OUT} ({os.path.getsize(OUT)} B) and {OUT_JSON} ({os.path.getsize(OUT_JSON)} B)")
