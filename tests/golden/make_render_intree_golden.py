"""Golden vectors from the reference's IN-TREE renderer code, executed UNMODIFIED from /root/reference against the
pytorch3d stand-in under ``oracle/pt3d_stub`` (pytorch3d itself is not installable here).  Run in the build container:

    python tests/golden/make_render_intree_golden.py        # writes tests/golden/render_intree_ref.npz

What runs is the reference's own ``MLPWithInputSkips`` (custom_modules.py:44-160), ``RenderMLP.forward`` /
``get_normals`` (holo_voxel_grid_implicit_function.py:48-145), ``HoloVoxelGridImplicitFunction.forward`` (:182-269)
and ``HoloMultiPassEmissionAbsorptionRenderer._run_raymarcher`` (holo_multipass_ea.py:79-125).  The vectors pin the
oracle's restatement of THAT logic (and the RenderMLP state-dict key names).  The pytorch3d leaves underneath are the
stub's (= the oracle's restatement), so the leaf arithmetic stays unpinned -- see oracle/pt3d_stub/README.md.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
OUT = os.path.join(HERE, "render_intree_ref.npz")


def generate():
    for p in (ROOT, REF, os.path.join(ROOT, "oracle", "pt3d_stub")):   # REF ahead of ROOT: holo_diffusion = the reference, not the shim
        if p not in sys.path:
            sys.path.insert(0, p)
    from holo_diffusion.holo_multipass_ea import HoloMultiPassEmissionAbsorptionRenderer
    from holo_diffusion.holo_voxel_grid_implicit_function import HoloVoxelGridImplicitFunction, RenderMLP
    from pytorch3d.implicitron.models.renderer.base import EvaluationMode, ImplicitronRayBundle

    from oracle import render_oracle as ro

    out = {}
    C, R, EXT, F_HEAD = 16, 8, 8.0, 4
    g = torch.Generator().manual_seed(11)

    def rnd(*s):
        return torch.randn(*s, generator=g)

    # ---- the implicit function with the reference's own RenderMLP inside (feature head on: feature_dim = 4)
    torch.manual_seed(7)
    fn = HoloVoxelGridImplicitFunction(resol=R, volume_extent=EXT, n_hidden=C, feature_dim=F_HEAD, render_normals=True,
                                       render_mlp_args=dict(dnet_hidden_dim=64))   # 64 hidden units keep the file small
    with torch.no_grad():   # make compositing non-trivial: a positive, varied density
        last = fn.render_mlp._density_net.mlp[-1][0]
        last.weight[-1] *= 8.0
        last.bias[-1] += 0.5
    sd = {k: v.detach().clone() for k, v in fn.render_mlp.state_dict().items()}
    for k, v in sd.items():
        out["sd/" + k] = v.numpy()
    out["sd_keys"] = np.array(list(sd.keys()))
    # ---- RenderMLP.forward on raw features / directions
    feats, dirs = torch.tanh(rnd(37, C)), torch.nn.functional.normalize(rnd(37, 3), dim=-1)
    with torch.no_grad():
        d, rgb, head = fn.render_mlp(feats, dirs)
    out.update({"mlp/feats": feats.numpy(), "mlp/dirs": dirs.numpy(), "mlp/dens": d.numpy(), "mlp/rgb": rgb.numpy(),
                "mlp/head": head.numpy()})
    # ---- HoloVoxelGridImplicitFunction.forward: ray bundle (with normals) and explicit pts_3d
    grid = torch.tanh(rnd(1, C, R, R, R))
    cams = ro.simple_360_cameras(8)
    b = ro.sample_rays(cams[2], 6, 5, 7)
    bundle = ImplicitronRayBundle(b.origins, b.directions, b.lengths, b.xys)
    dens, f, aux = fn(ray_bundle=bundle, voxel_grid_features=grid)
    out.update({"if/grid": grid.numpy(), "if/origins": b.origins.numpy(), "if/directions": b.directions.numpy(),
                "if/lengths": b.lengths.numpy(), "if/xys": b.xys.numpy(), "if/dens": dens.detach().numpy(),
                "if/feats": f.detach().numpy(), "if/normals": aux["normals"].detach().numpy()})
    pts = (rnd(2, 9, 3) * 2.5)
    fn.render_normals = False
    with torch.no_grad():
        dens_p, f_p, aux_p = fn(pts_3d=pts, voxel_grid_features=grid)
    assert aux_p == {}
    out.update({"pts/pts": pts.numpy(), "pts/dens": dens_p.numpy(), "pts/feats": f_p.numpy()})

    # ---- the multi-pass renderer: the reference's _run_raymarcher recursion
    class Bound:   # what pytorch3d's ImplicitFunctionWrapper.bind_args does for voxel_grid_features
        def __init__(self, f_, **kw):
            self.f, self.kw = f_, kw

        def __call__(self, **kw):
            return self.f(**kw, **self.kw)

    def stages(o, tag):
        i = 0
        while o is not None:
            out[f"{tag}/stage{i}/features"] = o.features.detach().numpy()
            out[f"{tag}/stage{i}/depths"] = o.depths.detach().numpy()
            out[f"{tag}/stage{i}/masks"] = o.masks.detach().numpy()
            out[f"{tag}/stage{i}/has_weights"] = np.array(o.weights is not None)
            if o.weights is not None:
                out[f"{tag}/stage{i}/weights"] = o.weights.detach().numpy()
            out[f"{tag}/stage{i}/has_normals"] = np.array(o.normals is not None)
            if o.normals is not None:
                out[f"{tag}/stage{i}/normals"] = o.normals.detach().numpy()
            o = o.prev_stage
            i += 1
        out[f"{tag}/n_stages"] = np.array(i)

    N_FINE = 4
    bg = (1.0, 1.0, 1.0)
    kw = dict(n_pts_per_ray_fine_evaluation=N_FINE, n_pts_per_ray_fine_training=N_FINE,
              raymarcher_EmissionAbsorptionRaymarcher_args=dict(bg_color=bg))
    fn.render_normals = True
    fn.feature_dim = 0   # the model forces the colour-only head (holo_diffusion_model.py:156); rebuild the MLP head-less
    fn.render_mlp._feature_net = None
    ifs = [Bound(fn, voxel_grid_features=grid), Bound(fn, voxel_grid_features=grid)]
    r = HoloMultiPassEmissionAbsorptionRenderer(return_weights=True, **kw)
    with torch.no_grad():   # as generate_samples.py runs it; get_normals re-enables grad inside (:139)
        stages(r(bundle, ifs, EvaluationMode.EVALUATION), "eval_w")
    fn.render_normals = False
    r2 = HoloMultiPassEmissionAbsorptionRenderer(return_weights=False, **kw)
    with torch.no_grad():
        stages(r2(bundle, ifs, EvaluationMode.EVALUATION), "eval_now")
        # training mode: density noise (std forced to 1.0 by the Holo subclass) + stratified refinement, seeded
        assert r.density_noise_std_train == 1.0
        torch.manual_seed(123)
        stages(r(bundle, ifs, EvaluationMode.TRAINING), "train_w")
    out["meta/C_R_EXT_NFINE"] = np.array([C, R, EXT, N_FINE], dtype=np.float64)
    return out


if __name__ == "__main__":
    o = generate()
    np.savez_compressed(OUT, **o)
    print(f"wrote {OUT}: {len(o)} arrays, {os.path.getsize(OUT)} bytes")
