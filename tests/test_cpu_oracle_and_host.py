"""CPU tier (`-m "not gpu"`): the oracle against the committed reference vectors, the host-side mirror of the
reference interface, and the C-ABI surface (symbols only -- no compute without a GPU)."""
import ctypes
import json
import math
import os

import numpy as np
import pytest
import torch

from conftest import rel_err
from fixtures import make_grid, make_mlp
from oracle import diffusion_oracle as do
from oracle import render_oracle as ro
from oracle import unet_oracle as uo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {
    "base16": dict(in_ch=16, R=16, model_ch=64, num_res_blocks=2, channel_mult=(1, 1, 2, 4, 8), attention_resolutions=(4, 8), heads=2, seed=2),
    "small8": dict(in_ch=8, R=8, model_ch=32, num_res_blocks=1, channel_mult=(1, 2), attention_resolutions=(2,), heads=1, seed=5),
}


# ----------------------------------------------------------------------------- oracle vs reference vectors
@pytest.mark.parametrize("name", ["small8", "base16"])
@pytest.mark.parametrize("t", [0, 500])
def test_unet_oracle_matches_reference_vectors(name, t):
    c = CASES[name]
    g = np.load(os.path.join(GOLD, "unet_ref.npz"))
    sd = uo.make_unet_state_dict(c["in_ch"], c["in_ch"], c["model_ch"], c["num_res_blocks"], c["channel_mult"],
                                 c["attention_resolutions"], seed=c["seed"])
    x = make_grid(c["in_ch"], c["R"], seed=0)
    with torch.no_grad():
        y = uo.unet_forward(sd, x, torch.full((1,), t, dtype=torch.long), n_heads=c["heads"])
    ref = torch.from_numpy(g[f"{name}_t{t}"])
    assert rel_err(y.reshape(-1)[::4], ref) < 2e-6  # same ops, same order (bit-identical where it was generated)
    st = g[f"{name}_t{t}_stats"]
    assert abs(y.abs().max().item() - st[2]) < 1e-4 * st[2]


def test_diffusion_oracle_matches_reference_vectors():
    g = np.load(os.path.join(GOLD, "diffusion_ref.npz"))
    g2 = np.load(os.path.join(GOLD, "ddim_ref.npz"))
    tab = do.schedule_tables()
    for k in tab:
        assert np.array_equal(tab[k], g[k] if k in g else g2[k]), k  # fp64 tables, bit-exact
    x, noise, t = torch.from_numpy(g["x"]), torch.from_numpy(g["noise"]), torch.from_numpy(g["t"])
    out = do.p_sample(tab, lambda z, tt: torch.tanh(1.7 * z) * 1.3, x, t, noise)
    assert torch.equal(out["sample"], torch.from_numpy(g["p_sample"]))
    assert torch.equal(out["pred_xstart"], torch.from_numpy(g["pred_xstart"]))
    assert torch.equal(do.q_sample(tab, x, t, noise), torch.from_numpy(g["q_sample"]))
    # DDIM (forward eta 0 / 0.5, reverse ODE) against the reference's ddim_sample / ddim_reverse_sample
    x, noise, t = torch.from_numpy(g2["x"]), torch.from_numpy(g2["noise"]), torch.from_numpy(g2["t"])
    model = lambda z, tt: torch.tanh(1.7 * z) * 1.3  # noqa: E731
    for eta in (0.0, 0.5):
        assert torch.equal(do.ddim_sample(tab, model, x, t, noise, eta=eta)["sample"], torch.from_numpy(g2[f"ddim_eta{eta}"]))
    rev = do.ddim_sample(tab, model, x, t, noise, reverse=True)
    assert torch.equal(rev["sample"], torch.from_numpy(g2["ddim_reverse"]))
    assert torch.equal(rev["pred_xstart"], torch.from_numpy(g2["pred_xstart"]))


def test_render_oracle_regression_and_properties():
    g = np.load(os.path.join(GOLD, "render_golden.npz"))
    C, R, HW, S = 8, 8, 12, 8
    grid, p = make_grid(C, R), make_mlp(C)
    cams = ro.simple_360_cameras(8)
    assert rel_err(cams.R, torch.from_numpy(g["R"])) < 1e-6
    b = ro.sample_rays(cams[3], HW, HW, S)
    o = ro.render_chunked(p, grid, b, R, 8.0, 2, 4, chunk_size_grid=0)
    assert rel_err(o.features, torch.from_numpy(g["features"])) < 1e-5
    assert rel_err(o.prev_stage.weights, torch.from_numpy(g["prev_weights"])) < 1e-5
    # chunked == unchunked, bit for bit (GenericModel._render re-assembly)
    o2 = ro.render_chunked(p, grid, b, R, 8.0, 2, 4, chunk_size_grid=64)
    assert torch.equal(o.features, o2.features) and torch.equal(o.lengths, o2.lengths)
    # emission-absorption weights are a sub-probability distribution; masks = their sum when the last sample
    # absorbs the rest
    w = o.weights
    assert float(w.min()) >= 0 and float(w.sum(-1).max()) <= 1 + 1e-5
    # refined depths are sorted and contain the coarse depths
    assert bool((o.lengths[..., 1:] >= o.lengths[..., :-1]).all())
    # camera rig: rotations orthonormal, camera centres at radius 10, rays unit length, ray order row-major
    assert rel_err(cams.R @ cams.R.transpose(1, 2), torch.eye(3).expand(8, 3, 3)) < 1e-5
    assert rel_err(cams.centre().norm(dim=-1), torch.full((8,), 10.0)) < 1e-5
    assert rel_err(b.directions.norm(dim=-1), torch.ones(1, HW, HW)) < 1e-5
    xy = ro.ndc_xy_grid(HW, HW)
    assert torch.equal(b.xys[0], xy) and xy[0, 0, 0] > xy[0, -1, 0] and xy[0, 0, 1] > xy[-1, 0, 1]


def test_collapsed_mlp_equals_layered_oracle():
    """The algebra behind holo_affine_compose_f64: the activation-free density layers collapse to one affine map."""
    C = 16
    p = make_mlp(C)
    A = p["_density_net.mlp.0.0.weight"].double()
    c = p["_density_net.mlp.0.0.bias"].double()
    for li in (1, 2, 3):
        W, b = p[f"_density_net.mlp.{li}.0.weight"].double(), p[f"_density_net.mlp.{li}.0.bias"].double()
        rows = A.shape[0]
        A_new = W[:, :rows] @ A + (W[:, rows:] if li == 2 else 0)
        c = W[:, :rows] @ c + b
        A = A_new
    x = torch.randn(64, C)
    d = torch.nn.functional.normalize(torch.randn(64, 3), dim=-1)
    dens, _ = ro.render_mlp(p, x, d)
    coll = torch.nn.functional.leaky_relu(x.double() @ A.t() + c, 0.2)[:, -1:]
    assert rel_err(coll, dens) < 1e-5


# ----------------------------------------------------------------------------- C-ABI surface
def test_abi_exports_every_declared_symbol():
    from holo_diffusion_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 22 and "holo_render_fwd" in protos and "holo_conv3d_tc" in protos
    assert os.path.exists(_lib.LIB_PATH), "libholo_b200.so must be built in-tree (python holo_diffusion_b200/build.py)"
    cdll = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(cdll, name), f"libholo_b200.so does not export {name}"
    L = _lib.lib()
    assert L.cdll.holo_version() == L._header_version()
    assert L.cdll.holo_render_mlp_packed_floats(256, 32, 27) == (257 * 32 + 4 * 256 + 4 + 81 + 3 + 3) // 4 * 4


def test_no_cpu_fallback():
    import holo_diffusion_b200 as hd
    net = hd.SimpleUnet3D(image_size=8, in_channels=8, out_channels=8, model_channels=32, num_res_blocks=1,
                          channel_mult=(1, 2), attention_resolutions=(2,), num_heads=1)
    with pytest.raises(hd.HoloError):
        net(torch.zeros(1, 8, 8, 8, 8), torch.zeros(1, dtype=torch.long))


def test_product_never_imports_oracle():
    pkg = os.path.join(os.path.dirname(GOLD), "..", "holo_diffusion_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "import oracle" not in src and "from oracle" not in src, f


# ----------------------------------------------------------------------------- host mirror of the reference interface
def test_state_dict_keys_match_reference():
    import holo_diffusion_b200 as hd
    keys = json.load(open(os.path.join(GOLD, "unet_keys.json")))
    for name, c in CASES.items():
        net = hd.SimpleUnet3D(image_size=c["R"], in_channels=c["in_ch"], out_channels=c["in_ch"], model_channels=c["model_ch"],
                              num_res_blocks=c["num_res_blocks"], channel_mult=c["channel_mult"],
                              attention_resolutions=c["attention_resolutions"], num_heads=c["heads"])
        mine = {k: list(v.shape) for k, v in net._net.state_dict().items()}
        assert mine == keys[name]
        # SimpleUnet3D init: Conv3d / Linear biases are zero, attention proj_out is zero (diffusion_utils.py:77-80, unet.py:392)
        assert float(net._net.input_blocks[0][0].bias.abs().max()) == 0.0


def test_host_cameras_match_oracle():
    import holo_diffusion_b200 as hd
    cams = hd.get_simple_360_camera_trajectory(2 * math.pi, 8, -math.pi / 6, 10.0, ro.CANONICAL_CO3D_UP_AXIS, 3.2)
    oc = ro.simple_360_cameras(8)
    assert rel_err(cams.R, oc.R) < 1e-6 and rel_err(cams.T, oc.T) < 1e-6
    assert rel_err(cams.get_camera_center(), oc.centre()) < 1e-6
    xy = hd.AdaptiveRaySampler.ndc_pixel_grid(6, 10)
    assert torch.equal(xy, ro.ndc_xy_grid(6, 10))  # bit-exact pixel bookkeeping, non-square included


def test_diffusion_tables_match_oracle():
    import holo_diffusion_b200 as hd
    d = hd.ImplicitronGaussianDiffusion()
    tab = do.schedule_tables()
    for k, v in d.tables64.items():
        assert np.array_equal(v, tab[k]), k
    with pytest.raises(NotImplementedError):
        hd.ImplicitronGaussianDiffusion(model_mean_type="EPSILON")


def test_attention_head_padding_is_exact():
    """Heads narrower than 64 channels run on the fused kernel zero-padded to 64 (unet.py executor): the re-packed
    qkv / proj weights with the TRUE ch^-1/2 softmax scale give the reference attention block exactly."""
    import math
    import torch.nn.functional as F
    from holo_diffusion_b200.unet import _pad_proj_heads, _pad_qkv_heads
    g = torch.Generator().manual_seed(5)
    heads, ch, chp, T = 2, 32, 64, 96
    C = heads * ch
    x = torch.randn(1, C, T, generator=g, dtype=torch.float64)
    wq, bq = torch.randn(3 * C, C, 1, generator=g, dtype=torch.float64) / 8, torch.randn(3 * C, generator=g, dtype=torch.float64)
    wp, bp = torch.randn(C, C, 1, generator=g, dtype=torch.float64) / 8, torch.randn(C, generator=g, dtype=torch.float64)

    def attn(qkv, n_ch, scale2):   # QKVAttentionLegacy.forward, unet.py:438-455, with an explicit logit scale
        q, k, v = qkv.reshape(heads, 3 * n_ch, T).split(n_ch, 1)
        w = torch.softmax(torch.einsum("bct,bcs->bts", q, k) * scale2, -1)
        return torch.einsum("bts,bcs->bct", w, v).reshape(1, -1, T)

    ref = F.conv1d(attn(F.conv1d(x, wq, bq), ch, 1 / math.sqrt(ch)), wp, bp)
    wq2, bq2 = _pad_qkv_heads(heads, ch, chp)(wq, bq)
    wp2, bp2 = _pad_proj_heads(heads, ch, chp)(wp, bp)
    assert wq2.shape == (heads * 3 * chp, C, 1) and wp2.shape == (C, heads * chp, 1)
    got = F.conv1d(attn(F.conv1d(x, wq2, bq2), chp, 1 / math.sqrt(ch)), wp2, bp2)
    assert torch.allclose(got, ref, rtol=0, atol=1e-12)


# ------------------------------------------------------------------------------------------------------------
# the reference's IN-TREE renderer code (executed unmodified against oracle/pt3d_stub) vs the oracle restatement
# ------------------------------------------------------------------------------------------------------------
def _intree():
    g = np.load(os.path.join(GOLD, "render_intree_ref.npz"))
    sd = {str(k): torch.from_numpy(g["sd/" + str(k)]) for k in g["sd_keys"]}
    return g, sd


def _t(g, k):
    return torch.from_numpy(g[k])


def test_render_oracle_matches_reference_intree_code():
    """RenderMLP.forward / get_normals, HoloVoxelGridImplicitFunction.forward (ray bundle and pts_3d) and the
    HoloMultiPassEmissionAbsorptionRenderer recursion: outputs of the reference's own code (tests/golden/
    make_render_intree_golden.py) against oracle.render_oracle.  Same torch ops in the same order => tight bounds.
    Pins the in-tree logic; the pytorch3d leaves underneath remain unpinned (oracle/pt3d_stub/README.md)."""
    from oracle import render_oracle as ro
    g, sd = _intree()
    C, R, EXT, NF = (float(v) for v in g["meta/C_R_EXT_NFINE"])
    R, NF = int(R), int(NF)
    # state-dict names: the product's RenderMLP uses the same keys as the reference module
    from holo_diffusion_b200.renderer import RenderMLP
    ours = RenderMLP(input_dims=int(C), output_vp_independent_feature_dims=4, dnet_hidden_dim=64)
    assert list(ours.state_dict().keys()) == [str(k) for k in g["sd_keys"]]
    assert all(tuple(v.shape) == tuple(sd[k].shape) for k, v in ours.state_dict().items())
    # RenderMLP.forward incl. the view-independent head
    d, rgb, head = ro.render_mlp(sd, _t(g, "mlp/feats"), _t(g, "mlp/dirs"), return_head=True)
    for got, key in ((d, "mlp/dens"), (rgb, "mlp/rgb"), (head, "mlp/head")):
        assert torch.allclose(got, _t(g, key), rtol=0, atol=1e-6), key
    # implicit function on a ray bundle, with normals
    grid = _t(g, "if/grid")
    b = ro.OracleRayBundle(_t(g, "if/origins"), _t(g, "if/directions"), _t(g, "if/lengths"), _t(g, "if/xys"))
    dens, feats, normals = ro.implicit_function(sd, grid, b, R, EXT, render_normals=True)
    assert torch.allclose(dens, _t(g, "if/dens"), rtol=0, atol=1e-6)
    assert torch.allclose(feats, _t(g, "if/feats"), rtol=0, atol=1e-6) and feats.shape[-1] == 3 + 4
    assert torch.allclose(normals, _t(g, "if/normals"), rtol=0, atol=1e-5)
    # explicit points (dummy all-ones directions, :229-237)
    dens_p, feats_p = ro.implicit_function(sd, grid, None, R, EXT, pts_3d=_t(g, "pts/pts"))
    assert torch.allclose(dens_p, _t(g, "pts/dens"), rtol=0, atol=1e-6)
    assert torch.allclose(feats_p, _t(g, "pts/feats"), rtol=0, atol=1e-6)
    # multi-pass recursion (colour-only head, as the model forces): evaluation with weights + normals
    sd3 = {k: v for k, v in sd.items() if not k.startswith("_feature_net")}

    def check(o, tag, fields):
        i = 0
        while o is not None:
            for f in fields:
                assert torch.allclose(getattr(o, f), _t(g, f"{tag}/stage{i}/{f}"), rtol=0, atol=2e-6), (tag, i, f)
            o = o.prev_stage
            i += 1
        assert i == int(g[f"{tag}/n_stages"]) == 2

    o = ro.render_multipass(sd3, grid, b, R, EXT, 2, NF, (1.0, 1.0, 1.0), render_normals=True)
    check(o, "eval_w", ("features", "depths", "masks", "weights", "normals"))
    assert bool(g["eval_w/stage0/has_weights"]) and bool(g["eval_w/stage1/has_weights"])
    # return_weights=False: the reference drops the weights of EVERY stage (holo_multipass_ea.py:111-112), no normals
    o = ro.render_multipass(sd3, grid, b, R, EXT, 2, NF, (1.0, 1.0, 1.0))
    check(o, "eval_now", ("features", "depths", "masks"))
    assert not bool(g["eval_now/stage0/has_weights"]) and not bool(g["eval_now/stage1/has_weights"])
    assert not bool(g["eval_now/stage0/has_normals"])
    # training mode: density noise of std 1.0 (the Holo subclass's default, :76-77) per pass + stratified refinement;
    # draw order = noise(pass 0), uniforms(refiner), noise(pass 1)
    torch.manual_seed(123)
    S = b.lengths.shape[-1]
    n0 = torch.randn(*b.lengths.shape)
    u = torch.rand(*b.lengths.shape[:-1], NF)
    n1 = torch.randn(*b.lengths.shape[:-1], S + NF)
    o = ro.render_multipass(sd3, grid, b, R, EXT, 2, NF, (1.0, 1.0, 1.0), noise=[n0, n1], u=u)
    check(o, "train_w", ("features", "depths", "masks", "weights"))


@pytest.mark.skipif(not os.path.isdir("/root/reference/holo_diffusion"), reason="needs the reference checkout")
def test_intree_golden_is_reproducible_from_the_reference():
    """The committed vectors really are what the reference's code produces here (build container only)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_render_intree_golden",
                                                  os.path.join(GOLD, "make_render_intree_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    fresh = m.generate()
    g = np.load(os.path.join(GOLD, "render_intree_ref.npz"))
    assert sorted(fresh.keys()) == sorted(g.files)
    for k in g.files:
        a, b_ = np.asarray(fresh[k]), g[k]
        if a.dtype.kind in "fc":
            assert np.allclose(a, b_, rtol=0, atol=1e-6), k
        else:
            assert (a == b_).all(), k


def test_split_kv_merge_arithmetic_and_dispatch():
    """Split-KV fused attention (holo_attention_flash kv_splits > 1): each share of the keys is reduced against its OWN
    stabiliser m_z; the merge O = sum_z 2^(m_z - M) O_z / sum_z 2^(m_z - M) l_z (flash_combine_kernel) must reproduce
    the un-split softmax for ANY stabilisers.  Restated here in torch; plus the executor's choice of the split count."""
    import math
    from holo_diffusion_b200.unet import kv_split_for
    g = torch.Generator().manual_seed(9)
    T, ch, splits = 256, 16, 4
    q, k, v = (torch.randn(T, ch, generator=g, dtype=torch.float64) for _ in range(3))
    scale_log2 = (1 / math.sqrt(ch)) * math.log2(math.e)
    s = q @ k.t()
    ref = torch.softmax(s / math.sqrt(ch), -1) @ v
    per = T // splits
    parts = []
    for z in range(splits):
        sz = s[:, z * per:(z + 1) * per]
        m = sz.max(-1, keepdim=True).values + torch.rand(T, 1, generator=g, dtype=torch.float64)   # an inexact stabiliser
        p = torch.exp2(sz * scale_log2 - m * scale_log2)
        parts.append((m * scale_log2, p.sum(-1, keepdim=True), p @ v[z * per:(z + 1) * per]))
    M = torch.stack([m for m, _, _ in parts]).max(0).values
    num = sum(torch.exp2(m - M) * o for m, _, o in parts)
    den = sum(torch.exp2(m - M) * l for m, l, _ in parts)
    assert torch.allclose(num / den, ref, rtol=0, atol=1e-12)
    assert kv_split_for(4096, 2, "1") == 1 and kv_split_for(4096, 2, "3") == 3
    assert kv_split_for(4096, 2, "auto") == 2        # 64 CTAs -> 128
    assert kv_split_for(512, 2, "auto") == 4         # 8 CTAs, 8 key tiles -> 2 tiles per CTA
    assert kv_split_for(64, 1, "auto") == 1 and kv_split_for(32768, 2, "auto") == 1


# ------------------------------------------------------------------------------------------------------------
# independent anchors for the pytorch3d LEAVES of the renderer oracle (unpinned against pytorch3d itself): closed forms
# of volume rendering, independent implementations (scipy / numpy), geometric invariants
# ------------------------------------------------------------------------------------------------------------
def test_leaf_so3_exp_map_matches_scipy():
    from scipy.spatial.transform import Rotation
    g = torch.Generator().manual_seed(3)
    v = torch.randn(16, 3, generator=g, dtype=torch.float64) * 1.3
    got = ro.so3_exp_map(v)
    ref = torch.from_numpy(Rotation.from_rotvec(v.numpy()).as_matrix())
    assert torch.allclose(got, ref, atol=1e-12)


def test_leaf_camera_and_ray_geometry():
    """look-at cameras of the turntable: orthonormal R, centre at the orbit radius, the central pixel's ray goes
    through the scene centre, depths span [radius - extent, radius + extent], unit directions."""
    cams = ro.simple_360_cameras(8, dtype=torch.float64)
    R = cams.R
    eye = torch.eye(3, dtype=torch.float64)
    assert torch.allclose(R @ R.transpose(1, 2), eye.expand_as(R), atol=1e-12)
    assert torch.allclose(torch.linalg.det(R), torch.ones(8, dtype=torch.float64), atol=1e-12)
    C = cams.centre()
    assert torch.allclose(C.norm(dim=-1), torch.full((8,), 10.0, dtype=torch.float64), atol=1e-9)
    for pose in (0, 3, 6):
        b = ro.sample_rays(cams[pose], 5, 5, 9)                      # odd size: pixel (2, 2) sits at NDC (0, 0)
        o, d, z = b.origins[0, 2, 2], b.directions[0, 2, 2], b.lengths[0, 2, 2]
        assert torch.allclose(b.xys[0, 2, 2], torch.zeros(2, dtype=b.xys.dtype), atol=1e-7)
        assert abs(float(d.norm()) - 1.0) < 1e-6
        t_centre = 0.5 * (z[0] + z[-1])                              # depth of the scene centre along the ray
        assert float((o + t_centre * d).norm()) < 1e-4               # ... where the ray meets the origin
        assert abs(float(z[0]) - 6.0) < 1e-4 and abs(float(z[-1]) - 14.0) < 1e-4
        assert torch.allclose(z[1:] - z[:-1], torch.full((8,), 1.0, dtype=z.dtype), atol=1e-5)   # linspace


def test_leaf_trilinear_sampling_reproduces_linear_fields():
    """Volume convention: world (x, y, z) <-> grid axes (W, H, D), voxel centres at (i - (R-1)/2) * extent / R; trilinear
    interpolation is exact on a linear field; outside the volume the value is 0 (zeros padding)."""
    R, ext = 8, 8.0
    vs = ext / R
    c = (torch.arange(R, dtype=torch.float64) - (R - 1) / 2) * vs
    zz, yy, xx = torch.meshgrid(c, c, c, indexing="ij")                  # (D, H, W)
    coef = torch.tensor([[0.3, -1.1, 0.7, 0.2], [1.5, 0.4, -0.6, -0.9]], dtype=torch.float64)
    grid = torch.stack([k[0] * xx + k[1] * yy + k[2] * zz + k[3] for k in coef])[None]   # (1, 2, D, H, W)
    g = torch.Generator().manual_seed(4)
    p = (torch.rand(200, 3, generator=g, dtype=torch.float64) - 0.5) * (R - 1) * vs     # inside the centre lattice
    got = ro.sample_grid(grid, ro.world_to_local(p, R, ext))
    ref = p @ coef[:, :3].t() + coef[:, 3]
    assert torch.allclose(got, ref, atol=1e-10)
    far = torch.tensor([[10.0, 0.0, 0.0], [0.0, -9.0, 0.0]], dtype=torch.float64)
    assert float(ro.sample_grid(grid, ro.world_to_local(far, R, ext)).abs().max()) == 0.0


def test_leaf_emission_absorption_closed_forms():
    """NeRF volume rendering: constant density sigma over uniform steps delta gives w_i = (1 - e^{-sigma delta})
    e^{-sigma delta i}, the last interval is opaque (delta = 1e10) so the weights sum to 1; empty space returns the
    background colour, zero depth and zero mask."""
    S, sigma, delta = 12, 0.8, 0.5
    z = 2.0 + delta * torch.arange(S, dtype=torch.float64)[None]
    dens = torch.full((1, S, 1), sigma, dtype=torch.float64)
    col = torch.tensor([0.2, 0.5, 0.9], dtype=torch.float64).expand(1, S, 3)
    o = ro.ea_raymarch(dens, col, z, bg=(1.0, 1.0, 1.0))
    i = torch.arange(S, dtype=torch.float64)
    w = (1 - math.exp(-sigma * delta)) * torch.exp(-sigma * delta * i)
    w[-1] = math.exp(-sigma * delta * (S - 1))                           # opaque last interval
    assert torch.allclose(o.weights[0], w, atol=1e-12) and abs(float(o.weights.sum()) - 1.0) < 1e-12
    assert torch.allclose(o.features[0], col[0, 0], atol=1e-12) and abs(float(o.masks) - 1.0) < 1e-12
    assert abs(float(o.depths) - float((w * z[0]).sum())) < 1e-12
    e = ro.ea_raymarch(torch.zeros(1, S, 1, dtype=torch.float64), col, z, bg=(0.1, 0.2, 0.3))
    assert torch.allclose(e.features[0], torch.tensor([0.1, 0.2, 0.3], dtype=torch.float64), atol=1e-12)
    assert float(e.masks) == 0.0 and float(e.depths) == 0.0 and float(e.weights.abs().max()) == 0.0
    neg = ro.ea_raymarch(torch.full((1, S, 1), -3.0, dtype=torch.float64), col, z)   # density_relu
    assert float(neg.masks) == 0.0


def test_leaf_sample_pdf_is_the_piecewise_linear_inverse_cdf():
    """sample_pdf (deterministic u) against numpy's interp of the same piecewise-linear CDF -- an independent
    implementation; the refiner then merges and sorts."""
    g = torch.Generator().manual_seed(6)
    n, S, N = 7, 10, 16
    bins = torch.sort(torch.rand(n, S, generator=g, dtype=torch.float64) * 4 + 2, -1)[0]
    w = torch.rand(n, S - 1, generator=g, dtype=torch.float64) + 0.05
    got = ro.sample_pdf(bins, w, N)
    wn = (w + 1e-5).numpy()
    cdf = np.concatenate([np.zeros((n, 1)), np.cumsum(wn / wn.sum(-1, keepdims=True), -1)], -1)
    u = np.linspace(0.0, 1.0, N)
    ref = np.stack([np.interp(u, cdf[r], bins[r].numpy()) for r in range(n)])
    assert np.allclose(got.numpy(), ref, atol=1e-9)
    # refiner: mids of the depths as bins, the inner weights, sorted union with the input samples
    z = bins
    wz = torch.rand(n, S, generator=g, dtype=torch.float64)
    zz = ro.refine_lengths(z, wz, N)
    assert zz.shape == (n, S + N) and bool((zz[:, 1:] >= zz[:, :-1]).all())
    mids = 0.5 * (z[:, 1:] + z[:, :-1])
    new = ro.sample_pdf(mids, wz[:, 1:-1], N)
    assert torch.allclose(torch.sort(torch.cat([z, new], -1), -1)[0], zz)


# ------------------------------------------------------------------------------------------------------------
# the reference's in-tree WRAPPERS (SimpleUnet3D, ImplicitronGaussianDiffusion, the turntable cameras), executed
# unmodified against oracle/pt3d_stub by tests/golden/make_wrappers_intree_golden.py
# ------------------------------------------------------------------------------------------------------------
def _wrappers():
    return np.load(os.path.join(GOLD, "wrappers_intree_ref.npz")), json.load(open(os.path.join(GOLD, "wrappers_intree_ref.json")))


def test_simple_unet3d_matches_reference_wrapper():
    """State-dict keys / shapes and the initialisation facts of the reference SimpleUnet3D (diffusion_utils.py:41-80):
    Xavier-uniform Conv3d / Linear weights with zero biases, Conv1d qkv left at PyTorch's default (non-zero bias),
    proj_out zeroed; and forward(x, t, cond_features) = UNet(cat(x, cond))."""
    from holo_diffusion_b200.unet import SimpleUnet3D
    g, facts = _wrappers()
    torch.manual_seed(0)
    net = SimpleUnet3D(image_size=16, in_channels=16, out_channels=16, model_channels=64, num_res_blocks=2,
                       channel_mult=(1, 1, 2, 4, 8), attention_resolutions=(4, 8), num_heads=2)
    ours = {k: list(v.shape) for k, v in net.state_dict().items()}
    assert ours == facts["unet_keys"]                                   # same names, same order, same shapes
    init = facts["unet_init"]
    mods = dict(net._net.named_modules())
    for name in init["zero_bias"]:
        assert float(mods[name].bias.detach().abs().max()) == 0.0, name
    for name in init["nonzero_bias"]:
        assert float(mods[name].bias.detach().abs().max()) > 0.0, name   # the Conv1d qkv layers keep their default init
    for name in init["zero_weight"]:
        assert float(mods[name].weight.detach().abs().max()) == 0.0, name  # zero_module(proj_out), unet.py:392
    for name in init["within_xavier_bound"]:
        w = mods[name].weight.detach()
        rf = int(np.prod(w.shape[2:])) if w.dim() > 2 else 1
        bound = math.sqrt(6.0 / ((w.shape[0] + w.shape[1]) * rf))
        assert 0.5 * bound < float(w.abs().max()) <= bound * (1 + 1e-6), name
    assert sorted(init["zero_bias"] + init["nonzero_bias"]) == sorted(
        n for n, m in mods.items() if isinstance(m, (torch.nn.Conv3d, torch.nn.Linear, torch.nn.Conv1d)))
    # cond_features: concatenated after x on the channel axis (oracle = the pinned UNet restatement)
    fix = uo.make_unet_state_dict(8, 8, 32, 1, (1, 2), (2,), seed=5)
    x, c = torch.from_numpy(g["unet/x"]), torch.from_numpy(g["unet/cond"])
    y = uo.unet_forward(fix, torch.cat([x, c], 1), torch.full((1,), 37, dtype=torch.long), n_heads=1)
    assert torch.allclose(y, torch.from_numpy(g["unet/y"]), rtol=0, atol=1e-5)


def test_diffusion_wrapper_defaults_match_reference():
    """ImplicitronGaussianDiffusion() with its DEFAULTS (diffusion_utils.py:89-112): linear betas 1e-4 .. 0.02, 1000
    steps, START_X / FIXED_SMALL, no timestep rescaling -- the tables of the product's wrapper and of the oracle."""
    from holo_diffusion_b200.diffusion import ImplicitronGaussianDiffusion
    g, facts = _wrappers()
    assert facts["diffusion"] == {"num_timesteps": 1000, "model_mean_type": "START_X", "model_var_type": "FIXED_SMALL",
                                  "rescale_timesteps": False}
    d = ImplicitronGaussianDiffusion()
    assert d.num_timesteps == 1000
    for k, v in d.tables64.items():
        if "diffusion/" + k in g.files:
            assert np.allclose(v, g["diffusion/" + k], rtol=1e-14, atol=0), k
    t = do.schedule_tables()
    for k in ("posterior_mean_coef1", "posterior_mean_coef2", "posterior_log_variance_clipped", "sqrt_alphas_cumprod"):
        assert np.allclose(np.asarray(t[k], dtype=np.float64), g["diffusion/" + k], rtol=1e-14, atol=0), k


def test_turntable_cameras_match_reference_function():
    """get_simple_360_camera_trajectory (flyaround.py:301-350), its own source executed on the stand-in leaves: the
    azimuth / elevation conversion, R = R_plane @ R_lookat (NOT the other order), T unchanged, (n, 1) focal length."""
    import holo_diffusion_b200 as hd
    g, _ = _wrappers()
    for n, max_angle in ((8, 2 * math.pi), (5, math.pi)):
        tag = f"cams{n}"
        oc = ro.simple_360_cameras(n, max_angle=max_angle)
        assert torch.allclose(oc.R, torch.from_numpy(g[tag + "/R"]), atol=1e-6)
        assert torch.allclose(oc.T, torch.from_numpy(g[tag + "/T"]), atol=1e-6)
        pc = hd.get_simple_360_camera_trajectory(max_angle, n, -math.pi / 6, 10.0, ro.CANONICAL_CO3D_UP_AXIS, 3.2)
        assert torch.allclose(pc.R, torch.from_numpy(g[tag + "/R"]), atol=1e-6)
        assert torch.allclose(pc.T, torch.from_numpy(g[tag + "/T"]), atol=1e-6)
        assert tuple(g[tag + "/focal"].shape) == (n, 1)      # one focal length for both axes; ours stores it per axis
        assert torch.allclose(pc.focal_length, torch.from_numpy(g[tag + "/focal"]).expand(n, pc.focal_length.shape[1]))
        assert float(pc.principal_point.abs().max()) == 0.0 and float(np.abs(g[tag + "/pp"]).max()) == 0.0
        # the other multiplication order is a different set of cameras: the vectors do discriminate
        Rp = ro.so3_exp_map(torch.cross(torch.tensor((0.0, -1.0, 0.0)), torch.tensor(ro.CANONICAL_CO3D_UP_AXIS), dim=-1)[None])[0]
        R_look = torch.linalg.solve(Rp[None].expand(n, 3, 3), oc.R)
        assert float((torch.bmm(R_look, Rp[None].expand(n, 3, 3)) - oc.R).abs().max()) > 1e-2


@pytest.mark.skipif(not os.path.isdir("/root/reference/holo_diffusion"), reason="needs the reference checkout")
def test_wrappers_golden_is_reproducible_from_the_reference():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_wrappers_intree_golden",
                                                  os.path.join(GOLD, "make_wrappers_intree_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    arrays, facts = m.generate()
    g, f = _wrappers()
    assert facts == f and sorted(arrays.keys()) == sorted(g.files)
    for k in g.files:
        assert np.allclose(np.asarray(arrays[k]), g[k], rtol=0, atol=1e-6), k


def test_plugin_signatures_match_reference_source():
    """The call signatures of the plug-in surface, read from the reference SOURCE by
    tests/golden/make_wrappers_intree_golden.py: same parameter names, order and kinds (keyword-only forward()s), same
    defaults where the reference states one; extra parameters of ours must be optional and keyword-only."""
    import inspect

    import holo_diffusion_b200 as hd
    from holo_diffusion_b200 import renderer as rd
    from holo_diffusion_b200 import unet as un
    _, facts = _wrappers()
    ours = {"HoloDiffusionModel.forward": hd.HoloDiffusionModel.forward,
            "HoloVoxelGridImplicitFunction.forward": hd.HoloVoxelGridImplicitFunction.forward,
            "RenderMLP.forward": rd.RenderMLP.forward,
            "HoloMultiPassEmissionAbsorptionRenderer._run_raymarcher": hd.HoloMultiPassEmissionAbsorptionRenderer._run_raymarcher,
            "Unet3DBase.forward": un.Unet3DBase.forward, "SimpleUnet3D.forward": un.SimpleUnet3D.forward,
            "get_simple_360_camera_trajectory": hd.get_simple_360_camera_trajectory}
    assert sorted(ours) == sorted(facts["signatures"])
    for name, fn in ours.items():
        fn = inspect.unwrap(fn)                       # forward() is wrapped by torch.no_grad()
        params = inspect.signature(fn).parameters
        ref = [(n, k) for n, k, _ in facts["signatures"][name]]
        got = [(p.name, p.kind.name) for p in params.values() if p.name in dict(ref)]
        assert got == ref, (name, got, ref)
        # anything extra must be optional and keyword-only (an extension the reference's callers never see)
        for p in params.values():
            if p.name not in dict(ref):
                assert p.kind.name == "KEYWORD_ONLY" and p.default is not inspect.Parameter.empty, (name, p.name)
        for n, _, default in facts["signatures"][name]:
            p = inspect.signature(fn).parameters.get(n)
            if default is None or p is None:
                continue
            if default == "None":
                assert p.default is None, (name, n)
            elif default == "EvaluationMode.EVALUATION":
                assert p.default == hd.EvaluationMode.EVALUATION, (name, n)
            else:
                assert p.default == eval(default), (name, n)
    # the call generate_samples.py makes (flyaround.py:247-253): a FrameData unpacked into keywords, image_rgb = None
    sig = inspect.signature(inspect.unwrap(hd.HoloDiffusionModel.forward))
    sig.bind(None, frame_number=None, sequence_category=None, image_rgb=None, camera=object(), fg_probability=None,
             mask_crop=None, depth_map=None, sequence_name=None, frame_timestamp=None,
             evaluation_mode=hd.EvaluationMode.EVALUATION, voxel_features=None)


def test_model_forward_orchestration_matches_reference_source():
    """HoloDiffusionModel.forward (holo_diffusion_model.py:201-540), its source executed on a stand-in `self` wired to
    the oracle (tests/golden/make_model_forward_intree_golden.py): the composition tanh(UNet(g, t = 0)) -> render of
    camera[0] only -> NCHW permutes reproduces its images; call order and preds keys as recorded."""
    g = np.load(os.path.join(GOLD, "model_forward_intree_ref.npz"))
    C, R, HW, S, NF = 8, 8, 6, 8, 4
    assert [str(x) for x in g["log"]] == ["('net_3d', [0])", "'bind'", "'bind'", "('rays', 1, 'EVALUATION', True)",
                                          "('render', 'FULL_GRID', 'EVALUATION', 2)", "'unbind'", "'unbind'"]
    assert bool(g["range_assert"])
    ref_keys = [str(k) for k in g["preds_keys"]]
    assert ref_keys == ["depths_render", "images_render", "implicitron_render", "masks_render", "ray_bundle", "rendered"]
    # the oracle pipeline the GPU tests compare the product with (tests/test_model_gpu.py) == the reference's forward
    sd = uo.make_unet_state_dict(C, C, 32, 1, (1, 2), (2,), seed=5)
    mlp = make_mlp(C)
    grid = torch.from_numpy(g["grid"])
    cams = ro.OracleCameras(*(torch.from_numpy(g[k]) for k in ("cam_R", "cam_T", "cam_focal", "cam_pp")))
    vox = torch.tanh(uo.unet_forward(sd, grid, torch.zeros(1, dtype=torch.long), n_heads=1))
    out = ro.render_chunked(mlp, vox, ro.sample_rays(cams[0], HW, HW, S), R, 8.0, 2, NF, chunk_size_grid=0)
    assert torch.allclose(out.features.permute(0, 3, 1, 2), torch.from_numpy(g["images_render"]), atol=1e-6)
    assert torch.allclose(out.depths.permute(0, 3, 1, 2), torch.from_numpy(g["depths_render"]), atol=1e-5)
    assert torch.allclose(out.masks.permute(0, 3, 1, 2), torch.from_numpy(g["masks_render"]), atol=1e-6)
    assert tuple(g["images_render"].shape) == (1, 3, HW, HW)            # one target view out of the 3 cameras
    # the product's forward returns these keys (plus its own extras), and no "objective" without losses
    import inspect

    import holo_diffusion_b200 as hd
    body = inspect.getsource(hd.HoloDiffusionModel.forward)
    for k in ref_keys:
        assert f'"{k}"' in body, k
    assert 'preds["objective"]' not in body


@pytest.mark.skipif(not os.path.isdir("/root/reference/holo_diffusion"), reason="needs the reference checkout")
def test_model_forward_golden_is_reproducible_from_the_reference():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_model_forward_intree_golden",
                                                  os.path.join(GOLD, "make_model_forward_intree_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    fresh = m.generate()
    g = np.load(os.path.join(GOLD, "model_forward_intree_ref.npz"))
    assert sorted(fresh.keys()) == sorted(g.files)
    for k in g.files:
        a, b_ = np.asarray(fresh[k]), g[k]
        if a.dtype.kind == "f":
            assert np.allclose(a, b_, rtol=0, atol=1e-6), k
        else:
            assert (a == b_).all(), k
