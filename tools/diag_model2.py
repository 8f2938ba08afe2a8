import sys, torch
sys.path.insert(0, "efficient-rpe-vit_b200"); sys.path.insert(0, "tests")
from conftest import load_golden, rel_l2
from erv_b200 import MNIST_CONFIG, create_model, _capi as C, ops
name = sys.argv[1] if len(sys.argv) > 1 else "performer_relu_most_general"
g = load_golden(f"model_{name}.npz")
caps = {}
for mode in (1, 0):
    C.load().erv_block_set_tensor_core(mode)
    model = create_model(name, MNIST_CONFIG, dropout=0.0)
    model.load_state_dict({k[3:]: v for k, v in g.items() if k.startswith("sd.")})
    model = model.to("cuda").eval()
    cap = {}
    orig = {}
    for fn in ("block_ln_qkv", "block_mlp"):
        orig[fn] = getattr(ops, fn)
    cnt = {"q": 0, "m": 0}
    def wrap_q(x, *a, **k):
        i = cnt["q"]; cnt["q"] += 1
        out = orig["block_ln_qkv"](x, *a, **k)
        out.register_hook(lambda gr, i=i: cap.__setitem__(f"dqkv{i}", gr.detach().clone()))
        x.register_hook(lambda gr, i=i: cap.__setitem__(f"dx_total{i}", gr.detach().clone()))
        cap[f"qkv{i}"] = out.detach().clone()
        return out
    def wrap_m(a_, x, *a, **k):
        i = cnt["m"]; cnt["m"] += 1
        a_.register_hook(lambda gr, i=i: cap.__setitem__(f"da{i}", gr.detach().clone()))
        out = orig["block_mlp"](a_, x, *a, **k)
        out.register_hook(lambda gr, i=i: cap.__setitem__(f"dy{i}", gr.detach().clone()))
        cap[f"y{i}"] = out.detach().clone(); cap[f"a{i}"] = a_.detach().clone()
        return out
    ops.block_ln_qkv, ops.block_mlp = wrap_q, wrap_m
    logits = model(g["images"].to("cuda"))
    torch.nn.functional.cross_entropy(logits, g["labels"].to("cuda")).backward()
    ops.block_ln_qkv, ops.block_mlp = orig["block_ln_qkv"], orig["block_mlp"]
    caps[mode] = cap
for k in sorted(caps[1]):
    a, b = caps[1][k], caps[0][k]
    e = rel_l2(a, b)
    rows = (a - b).reshape(-1, a.shape[-1]).norm(dim=1) / (b.reshape(-1, b.shape[-1]).norm(dim=1) + 1e-30)
    print("%-12s rel %.2e  shape %s  worst rows %s  max|b| %.3e" % (k, e, tuple(a.shape), rows.topk(3).indices.tolist(), b.abs().max()))
