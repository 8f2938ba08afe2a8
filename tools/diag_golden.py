"""Per-row error of one golden attention case on the GPU: python tools/diag_golden.py attn_d_relu_circulant_string.npz"""
import sys, os
sys.path.insert(0, 'efficient-rpe-vit_b200'); sys.path.insert(0, 'tests')
import torch
from conftest import load_golden, parse_attn_case, rel_l2, max_rel
from erv_b200 import ATTENTION_REGISTRY, RPE_REGISTRY
for fname in sys.argv[1:]:
    _, a, r = parse_attn_case(fname)
    g = load_golden(fname)
    heads = int(g["heads"]); _, n, dim = g["x"].shape
    kw = {"num_features": g["attn.omega"].shape[-1]} if a != "softmax" else {}
    attn = ATTENTION_REGISTRY[a](dim=dim, heads=heads, dropout=0.0, **kw)
    attn.load_state_dict({k[5:]: v for k, v in g.items() if k.startswith("attn.")})
    rpe = None
    if r is not None:
        rpe = RPE_REGISTRY[r](num_patches=n, dim=dim, heads=heads)
        rpe.load_state_dict({k[4:]: v for k, v in g.items() if k.startswith("rpe.")})
        rpe = rpe.cuda()
    attn = attn.cuda().eval()
    x = g["x"].cuda().requires_grad_(True)
    out = attn(x, rpe=rpe)
    (out * g["cotangent"].cuda()).sum().backward()
    print(fname, "TC2 disabled" if os.environ.get("ERV_DISABLE_TC2") else "TC2 enabled")
    print("  out rel_l2 %.3e max_rel %.3e | dx rel_l2 %.3e max_rel %.3e" % (rel_l2(out, g["out"]), max_rel(out, g["out"]), rel_l2(x.grad, g["dx"]), max_rel(x.grad, g["dx"])))
    e = (x.grad.cpu() - g["dx"]).norm(dim=-1); ref = g["dx"].norm(dim=-1)
    flat = e.flatten(); top = flat.topk(6)
    for v, i in zip(top.values, top.indices):
        b, t = divmod(int(i), n)
        print(f"    row b={b} n={t}: |err| {float(v):.3e} |ref| {float(ref[b, t]):.3e}  (rms row |ref| {float(ref.pow(2).mean().sqrt()):.3e})")
    for k, p in list(attn.named_parameters()) + (list(rpe.named_parameters()) if rpe is not None else []):
        key = ("grad.attn." + k) if ("grad.attn." + k) in g else ("grad.rpe." + k)
        print(f"  {k}: rel_l2 {rel_l2(p.grad, g[key]):.3e} max_rel {max_rel(p.grad, g[key]):.3e}")
