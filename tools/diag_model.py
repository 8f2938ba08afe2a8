import sys, torch
sys.path.insert(0, "efficient-rpe-vit_b200"); sys.path.insert(0, "tests")
from conftest import load_golden, rel_l2
from erv_b200 import MNIST_CONFIG, create_model, _capi as C, ops
name = sys.argv[1] if len(sys.argv) > 1 else "performer_relu_most_general"
g = load_golden(f"model_{name}.npz")
res = {}
for mode, fused in ((1, True), (0, True), (0, False)):
    C.load().erv_block_set_tensor_core(mode); ops.FUSED_BLOCK = fused
    model = create_model(name, MNIST_CONFIG, dropout=0.0)
    model.load_state_dict({k[3:]: v for k, v in g.items() if k.startswith("sd.")})
    model = model.to("cuda").eval()
    logits = model(g["images"].to("cuda"))
    loss = torch.nn.functional.cross_entropy(logits, g["labels"].to("cuda")); loss.backward()
    params = dict(model.named_parameters())
    res[(mode, fused)] = {k: rel_l2(params[k[5:]].grad, v) for k, v in g.items() if k.startswith("grad.")}
    res[(mode, fused)]["logits"] = rel_l2(logits, g["logits"])
keys = list(res[(1, True)])
print("%-60s %10s %10s %10s" % ("tensor", "tc", "ffma2", "unfused"))
for k in keys:
    print("%-60s %10.2e %10.2e %10.2e" % (k, res[(1, True)][k], res[(0, True)][k], res[(0, False)][k]))
