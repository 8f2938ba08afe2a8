"""Training-loop mechanics around the hot path: the captured step against torch.optim.Adam, and
checkpoint interchange in the reference's file layout (experiments/utils/training.py:373-443)."""
import copy

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _batches(n, bsz=8, seed=3):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(bsz, 1, 28, 28, generator=g).cuda(), torch.randint(0, 10, (bsz,), generator=g).cuda())
            for _ in range(n)]


def _model(name="performer_favor", seed=0):
    from erv_b200 import MNIST_CONFIG, create_model
    torch.manual_seed(seed)
    return create_model(name, MNIST_CONFIG, dropout=0.0).to("cuda").train()


def _torch_adam_steps(model, opt, batches):
    for img, lab in batches:
        opt.zero_grad()
        F.cross_entropy(model(img), lab).backward()
        opt.step()


@pytest.mark.parametrize("use_graph", [False, True])
def test_trainer_steps_match_torch_adam(use_graph):
    """N steps of the flat fused step (graph or eager) leave the parameters where torch.optim.Adam(lr=1e-3) leaves them;
    in particular the graph warm-up passes do not train."""
    from erv_b200.train import Trainer
    batches = _batches(4)
    ref = _model()
    ours = copy.deepcopy(ref)
    _torch_adam_steps(ref, torch.optim.Adam(ref.parameters(), lr=1e-3), batches)
    tr = Trainer(ours, lr=1e-3, use_graph=use_graph)
    for b in batches:
        tr.step(*b)
    assert int(tr.step_count) == len(batches)
    for (k, a), (_, b) in zip(ref.named_parameters(), ours.named_parameters()):
        assert rel_l2(b, a) < 2e-4, k  # Adam's first steps are sign-like: tiny gradient differences show up at lr scale


def test_checkpoint_from_torch_adam_resumes_in_trainer(tmp_path):
    """A checkpoint written the reference's way (model + torch.optim.Adam state) resumes in the Trainer, and the Trainer's
    checkpoint resumes under torch.optim.Adam: both continue like the uninterrupted torch run."""
    from erv_b200.train import Trainer, load_checkpoint, save_checkpoint
    batches = _batches(5)
    ref = _model("performer_relu_circulant")
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    _torch_adam_steps(ref, opt, batches[:2])
    path = str(tmp_path / "ref.pt")
    save_checkpoint(ref, opt, 7, {"val_acc": 12.5}, path)
    ck = torch.load(path)
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_state_dict", "metrics", "model_name", "attention_type",
                       "rpe_type"} and ck["model_name"] == "performer_relu_circulant"

    ours = _model("performer_relu_circulant", seed=99)  # different init and omega: everything must come from the file
    tr = Trainer(ours, lr=5e-2, use_graph=True)
    tr.step(*batches[0])  # a captured graph exists before the load
    epoch, metrics = load_checkpoint(ours, tr, path)
    assert (epoch, metrics) == (7, {"val_acc": 12.5}) and tr.lr == 1e-3 and int(tr.step_count) == 2
    assert torch.equal(ours.transformer_blocks[0].attention.omega, ref.transformer_blocks[0].attention.omega)
    for b in batches[2:4]:
        tr.step(*b)
    path2 = str(tmp_path / "ours.pt")
    save_checkpoint(ours, tr, 8, {}, path2)

    _torch_adam_steps(ref, opt, batches[2:4])  # the uninterrupted run
    for (k, a), (_, b) in zip(ref.named_parameters(), ours.named_parameters()):
        assert rel_l2(b, a) < 2e-4, k

    back = _model("performer_relu_circulant", seed=5)
    opt2 = torch.optim.Adam(back.parameters(), lr=0.5)
    load_checkpoint(back, opt2, path2)
    assert opt2.param_groups[0]["lr"] == 1e-3
    _torch_adam_steps(back, opt2, batches[4:])
    _torch_adam_steps(ref, opt, batches[4:])
    for (k, a), (_, b) in zip(ref.named_parameters(), back.named_parameters()):
        assert rel_l2(b, a) < 2e-4, k


def test_feature_redraw_interval_on_device():
    """favor_plus.py:168-171: in training mode omega is redrawn when redraw_counter % interval == 0 and the counter
    advances every call; eval mode touches neither.  The redrawn buffer stays on the module's device (the reference
    re-registers a CPU tensor)."""
    from erv_b200 import FAVORPlusAttention
    torch.manual_seed(0)
    att = FAVORPlusAttention(32, 2, num_features=32, feature_redraw_interval=2).cuda().train()
    x = torch.randn(2, 17, 32, device="cuda")
    seen = [att.omega.clone()]
    for _ in range(4):
        att(x)
        seen.append(att.omega.clone())
    changed = [not torch.equal(a, b) for a, b in zip(seen, seen[1:])]
    assert changed == [True, False, True, False] and int(att.redraw_counter) == 4 and att.omega.is_cuda
    g = att.omega[0].T @ att.omega[0]  # still orthogonal blocks scaled by sqrt(Dh)
    assert torch.allclose(g[:16, :16], 16 * torch.eye(16, device="cuda"), atol=1e-3)
    att.eval()
    att(x)
    assert int(att.redraw_counter) == 4 and torch.equal(att.omega, seen[-1])


def test_learning_rate_change_after_capture_is_honoured():
    """hyper-parameters used to be baked into the captured graph.  They now live in device memory: changing the
    learning rate after the first (captured) step, through param_groups, set_lr or a schedule, matches torch.optim.Adam."""
    from erv_b200.train import Trainer
    batches = _batches(6)
    ref = _model()
    ours = copy.deepcopy(ref)
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    _torch_adam_steps(ref, opt, batches[:2])
    for g in opt.param_groups:
        g["lr"] = 1e-2
    _torch_adam_steps(ref, opt, batches[2:4])
    for g in opt.param_groups:
        g["lr"] = 5e-4
    _torch_adam_steps(ref, opt, batches[4:])
    tr = Trainer(ours, lr=1e-3, use_graph=True)
    for b in batches[:2]:
        tr.step(*b)
    assert tr._graph is not None
    graph = tr._graph
    for g in tr.param_groups:
        g["lr"] = 1e-2
    assert tr.lr == 1e-2
    for b in batches[2:4]:
        tr.step(*b)
    tr.set_lr(5e-4)
    for b in batches[4:]:
        tr.step(*b)
    assert tr._graph is graph  # no re-capture was needed
    for (k, a), (_, b) in zip(ref.named_parameters(), ours.named_parameters()):
        assert rel_l2(b, a) < 3e-4, k
    # a schedule callable
    tr2 = Trainer(_model(), lr=1.0, use_graph=True)
    tr2.lr_schedule = lambda step: 1e-3 * (step + 1)
    for b in batches[:3]:
        tr2.step(*b)
    assert abs(tr2.lr - 3e-3) < 1e-12 and abs(float(tr2.hyper[0]) - 3e-3) < 1e-9


def test_feature_redraw_runs_outside_the_captured_step():
    """feature_redraw_interval used to sync (int(redraw_counter)) and redraw inside stream capture.  The Trainer
    now redraws on the host before the replay; the warm-up passes of the capture leave omega and the counter untouched."""
    from erv_b200 import MNIST_CONFIG, create_model
    from erv_b200.train import Trainer
    torch.manual_seed(0)
    model = create_model("performer_favor", MNIST_CONFIG, dropout=0.0,
                         attention_config={"feature_redraw_interval": 2}).to("cuda").train()
    attn = model.transformer_blocks[0].attention
    assert attn.feature_redraw_interval == 2
    tr = Trainer(model, lr=1e-3, use_graph=True)
    seen = []
    for i, b in enumerate(_batches(5)):
        loss = tr.step(*b)
        assert torch.isfinite(loss)
        seen.append(attn.omega.clone())
        assert int(attn.redraw_counter) == i + 1
    assert attn.feature_redraw_interval == 2
    assert torch.equal(seen[0], seen[1]) and torch.equal(seen[2], seen[3])  # steps 0, 2, 4 redraw
    assert not torch.equal(seen[1], seen[2]) and not torch.equal(seen[3], seen[4])
