"""holo_unet_create (csrc/unet_exec.cu) without a GPU: the C++ block list must reproduce the reference's module tree --
its parameter table has to equal the state-dict keys / element counts of the reference-named parameter tree
(holo_diffusion_b200.unet.UNetParams, itself pinned against the reference's SimpleUnet3D keys in
tests/golden/unet_keys.json) -- and the dry run must produce a workspace / packed-weights plan."""
import ctypes
import json
import os

import pytest

CFGS = [
    dict(in_channels=32, model_channels=64, num_res_blocks=2, channel_mult=(1, 1, 2, 4, 8), attention_resolutions=(4, 8), dims=(64, 64, 64)),
    dict(in_channels=16, model_channels=64, num_res_blocks=1, channel_mult=(1, 2), attention_resolutions=(1, 2), dims=(16, 16, 16)),
    dict(in_channels=32, model_channels=64, num_res_blocks=2, channel_mult=(1, 1, 2, 4, 8), attention_resolutions=(1, 2, 4, 8, 16),
         dims=(128, 128, 128)),
]


def _create(c):
    from holo_diffusion_b200 import ops
    from holo_diffusion_b200._lib import lib
    _i = ctypes.c_int
    cm, ar = list(c["channel_mult"]), list(c["attention_resolutions"])
    cfg = ops.HoloUnetConfig(c["in_channels"], c["model_channels"], c["in_channels"], c["num_res_blocks"], len(cm), (_i * 8)(*cm),
                             len(ar), (_i * 8)(*ar), 2, *c["dims"], 1, 1, 0, 1, 1)
    h = ctypes.c_void_p()
    lib().call("holo_unet_create", ctypes.byref(cfg), ctypes.byref(h))
    return h, lib()


@pytest.mark.parametrize("c", CFGS)
def test_parameter_table_matches_the_reference_names(c):
    from holo_diffusion_b200.unet import UNetParams
    h, L = _create(c)
    n = L.cdll.holo_unet_param_count(h)
    table = {L.cdll.holo_unet_param_name(h, i).decode(): L.cdll.holo_unet_param_numel(h, i) for i in range(n)}
    ref = UNetParams(c["in_channels"], c["model_channels"], c["in_channels"], c["num_res_blocks"], c["attention_resolutions"],
                     c["channel_mult"], 2).state_dict()
    assert table == {k: v.numel() for k, v in ref.items()}
    assert L.cdll.holo_unet_packed_bytes(h) > sum(table.values()) * 3        # pairs (2 x 2 B) + fp32 layouts / biases
    V = c["dims"][0] * c["dims"][1] * c["dims"][2]
    assert L.cdll.holo_unet_workspace_bytes(h) > V * c["model_channels"] * 4 * 4   # several full-resolution activations live
    L.cdll.holo_unet_destroy(h)


def test_base_args_names_are_the_reference_checkpoint_keys():
    """tests/golden/unet_keys.json["base16"] = names and shapes of the reference's own UNetModel (base args, 16 channels),
    written by tests/golden/make_golden.py from the imported reference."""
    import math
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "unet_keys.json")))["base16"]
    h, L = _create(dict(CFGS[0], in_channels=16))
    table = {L.cdll.holo_unet_param_name(h, i).decode(): L.cdll.holo_unet_param_numel(h, i)
             for i in range(L.cdll.holo_unet_param_count(h))}
    assert table == {k: math.prod(v) for k, v in ref.items()}


def test_set_param_rejects_unknown_names_and_wrong_sizes():
    from holo_diffusion_b200 import HoloError
    h, L = _create(CFGS[1])
    fake = ctypes.c_void_p(256)   # never dereferenced by set_param
    with pytest.raises(HoloError, match="unknown parameter"):
        L.call("holo_unet_set_param", h, b"input_blocks.99.0.weight", fake, 10)
    with pytest.raises(HoloError, match="elements"):
        L.call("holo_unet_set_param", h, b"time_embed.0.bias", fake, 7)
    with pytest.raises(HoloError, match="was not set"):
        L.call("holo_unet_pack", h, fake, None)
