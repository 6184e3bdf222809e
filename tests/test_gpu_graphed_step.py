"""parallel.GraphedTrainStep: the captured training step must train like the eager one."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


class _Model(torch.nn.Module):
    def __init__(self, G):
        super().__init__()
        self.backbone = G.GKGNet(choice="s", n_classes=80, size=192, drop_path=0.0)
        self.head = G.LabelQueryHead(80, 640)

    def forward(self, img, tgt):
        return sum(self.head.forward_train(self.backbone(img), tgt).values())


def test_graphed_step_matches_eager_training():
    import gkgnet_b200 as G
    from gkgnet_b200 import parallel as P
    G.set_norm_type("BN")
    torch.manual_seed(0)
    dev = torch.device("cuda")
    base = _Model(G).to(dev).train()
    g = torch.Generator(device="cuda").manual_seed(1)
    img = torch.randn(4, 3, 192, 192, device=dev, generator=g)
    tgt = (torch.rand(4, 80, device=dev, generator=g) < 0.1).float()

    # (1) lr = 0: the weights never move, so every replay must reproduce the eager step -- same loss, same clipped
    # gradients (up to the atomics' summation order and bf16)
    eager = copy.deepcopy(base)
    pe = [p for p in eager.parameters() if p.requires_grad]
    with torch.autocast("cuda", dtype=torch.bfloat16):
        want = eager(img, tgt)
    want.backward()
    torch.nn.utils.clip_grad_norm_(pe, 5.0)
    graphed = copy.deepcopy(base)
    pg = [p for p in graphed.parameters() if p.requires_grad]
    og = torch.optim.AdamW(pg, lr=0.0, weight_decay=0.0, fused=True, capturable=True)
    step = P.GraphedTrainStep(graphed, og, pg, img, tgt, clip_norm=5.0, warmup=2)
    assert step.launches_per_replay > 50
    for _ in range(2):
        got = step()
        assert abs(got.item() - want.item()) <= 1e-2 * abs(want.item()), (got.item(), want.item())
        ge = torch.cat([p.grad.flatten() for p in pe])
        gg = step.grads.flat
        cos = torch.nn.functional.cosine_similarity(ge, gg, dim=0).item()
        assert cos > 0.98, cos
        assert abs(gg.norm().item() - ge.norm().item()) <= 5e-2 * ge.norm().item()
    # running statistics advance inside the graph (one update per replay on top of warm-up + capture-free passes)
    bn = next(m for m in graphed.modules() if isinstance(m, torch.nn.BatchNorm2d))
    assert int(bn.num_batches_tracked) == 2 + 2

    # (2) lr > 0: replaying the graph trains, and new inputs go through the static buffers
    model = copy.deepcopy(base)
    pm = [p for p in model.parameters() if p.requires_grad]
    om = torch.optim.AdamW(pm, lr=2e-4, weight_decay=0.05, fused=True, capturable=True)
    step = P.GraphedTrainStep(model, om, pm, img, tgt, clip_norm=5.0, warmup=1)
    losses = [step().item() for _ in range(12)]
    assert sum(losses[-3:]) < sum(losses[:3]), losses
    img2 = torch.randn(4, 3, 192, 192, device=dev, generator=g)
    l2 = step(img2, tgt).item()
    assert l2 == l2


def test_prefetched_inputs_reach_the_graph():
    """prefetch() / step_prefetched(): the batch uploaded on the copy stream is the one the replay consumes."""
    import gkgnet_b200 as G
    from gkgnet_b200 import parallel as P
    G.set_norm_type("BN")
    torch.manual_seed(0)
    dev = torch.device("cuda")
    model = _Model(G).to(dev).train()
    pm = [p for p in model.parameters() if p.requires_grad]
    om = torch.optim.AdamW(pm, lr=0.0, weight_decay=0.0, fused=True, capturable=True)
    g = torch.Generator().manual_seed(4)
    batches = [(torch.randn(4, 3, 192, 192, generator=g).pin_memory(),
                (torch.rand(4, 80, generator=g) < 0.1).float().pin_memory()) for _ in range(3)]
    step = P.GraphedTrainStep(model, om, pm, batches[0][0].to(dev), batches[0][1].to(dev), clip_norm=5.0, warmup=1)
    direct = [step(i.to(dev), t.to(dev)).item() for i, t in batches]           # lr = 0: the loss depends on the batch only
    piped = []
    step.prefetch(*batches[0])
    for j in range(3):
        out = step.step_prefetched()
        if j + 1 < 3:
            step.prefetch(*batches[j + 1])
        piped.append(out.item())
    assert len(set(round(v, 3) for v in direct)) == 3                          # three different batches, three different losses
    for a, b in zip(direct, piped):
        assert abs(a - b) <= 5e-3 * abs(a), (direct, piped)
