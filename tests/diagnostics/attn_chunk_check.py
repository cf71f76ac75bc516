"""Fused attention against an fp64 softmax(QK^T)V on the same GPU for a list of (T, heads, ch, kv_splits) -- run in its
own process so that HOLO_ATTN_O_CHUNK (key tiles per O accumulation chain, read once by the library) can be varied:
    HOLO_ATTN_O_CHUNK=3 python tests/diagnostics/attn_chunk_check.py 1024,1,64,1 640,2,64,1 192,1,64,1
Prints one JSON object {case: rel_err}."""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from holo_diffusion_b200 import ops  # noqa: E402


def run(T, heads, ch, splits):
    g = torch.Generator().manual_seed(7)
    C = heads * ch
    qkv = (torch.randn(heads * 3 * ch, T, generator=g) * 1.2).cuda()
    q, k, v = qkv.reshape(heads, 3 * ch, T).double().split(ch, 1)
    ref = torch.empty(T, C, dtype=torch.float64, device="cuda")
    for h in range(heads):   # row blocks: T x T in fp64 does not fit for T = 32768 all at once on every box
        for r0 in range(0, T, 4096):
            w = torch.softmax(torch.einsum("ct,cs->ts", q[h][:, r0:r0 + 4096], k[h]) / math.sqrt(ch), -1)
            ref[r0:r0 + 4096, h * ch:(h + 1) * ch] = torch.einsum("ts,cs->tc", w, v[h])
    x = qkv.t().contiguous()
    hi = torch.empty(T, 3 * C, device="cuda", dtype=torch.float16)
    lo = torch.empty_like(hi)
    ops.split_bf16(x, T, 3 * C, 3 * C, hi, lo)
    vt_hi = torch.empty(C, T, device="cuda", dtype=torch.float16)
    vt_lo = torch.empty_like(vt_hi)
    ops.v_transpose_split(x, T, heads, ch, vt_hi, vt_lo)
    out = torch.full((T, C), float("nan"), device="cuda")
    ws = ops.attention_flash_workspace(T, heads, ch, splits, "cuda") if splits > 1 else None
    rc = ops.attention_flash(hi, lo, vt_hi, vt_lo, T, heads, ch, out, kv_splits=splits, workspace=ws)
    assert rc == 0
    torch.cuda.synchronize()
    return float((out.double() - ref).abs().max() / ref.abs().max())


if __name__ == "__main__":
    res = {"HOLO_ATTN_O_CHUNK": os.environ.get("HOLO_ATTN_O_CHUNK")}
    for case in sys.argv[1:]:
        T, heads, ch, splits = (int(x) for x in case.split(","))
        res[case] = run(T, heads, ch, splits)
    print(json.dumps(res))
