"""GPU: the label-query head kernels (scores, both losses, every gradient) against the CPU oracle's restatement of
LabelQueryHead.get_score / forward_train (label_query_head.py:49-85) with autograd.  fp32: 1e-4 of the scale."""
import pytest
import torch

from oracle import gkg_oracle as O

pytestmark = pytest.mark.gpu


def _close(a, b, tol=1e-4):
    return (a.float().cpu() - b.float().cpu()).abs().max().item() <= tol * max(1.0, b.float().abs().max().item())


@pytest.mark.parametrize("B,n,C", [(16, 80, 640), (3, 7, 48), (1, 80, 33)])
def test_label_head_scores_losses_and_gradients(B, n, C):
    import gkgnet_b200 as G
    torch.manual_seed(B + n)
    head = G.LabelQueryHead(n, C).cuda()
    with torch.no_grad():
        for p in head.parameters():
            p.normal_(0, 0.2)
    lab = torch.randn(B, n, C, device="cuda", requires_grad=True)
    gap = torch.randn(B, C, device="cuda", requires_grad=True)
    tgt = (torch.rand(B, n, device="cuda") < 0.3).float()
    score = head.get_score((lab, gap))
    losses = head.forward_train((lab, gap), tgt)
    total = sum(losses.values())
    total.backward()
    # oracle on the CPU with autograd
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in head.state_dict().items()}
    lc, gc = lab.detach().cpu().requires_grad_(True), gap.detach().cpu().requires_grad_(True)
    want_score = O.label_query_score(sd, lc, gc)
    want = O.head_losses(sd, lc, gc, tgt.cpu())
    sum(want.values()).backward()
    assert _close(score, want_score)
    for k in want:
        assert _close(losses[k], want[k]), k
    assert _close(lab.grad, lc.grad) and _close(gap.grad, gc.grad)
    for k, p in head.named_parameters():
        assert _close(p.grad, sd[k].grad), k


def test_asymmetric_loss_edge_cases():
    """Saturated scores (pt at the clip / eps boundaries) and gamma_pos != 0 against the host-side expression."""
    from gkgnet_b200 import ops
    from gkgnet_b200.head import asymmetric_loss
    s = torch.tensor([[-30.0, -3.0, -0.05, 0.0, 0.05, 3.0, 30.0, 2.9444]], device="cuda").repeat(2, 1).requires_grad_(True)
    t = torch.tensor([[0, 1, 0, 1, 0, 1, 0, 0], [1, 0, 1, 0, 1, 0, 1, 1]], device="cuda").float()
    for gp, gn, clip in ((0.0, 2.0, 0.05), (1.0, 4.0, 0.05), (0.0, 2.0, 0.0)):
        asl, bce = ops.multilabel_losses(s, t, gp, gn, clip, 1e-8, 0.1)
        (asl + bce).backward()
        got, s.grad = s.grad.clone(), None
        sc = s.detach().cpu().requires_grad_(True)
        ref = asymmetric_loss(sc, t.cpu(), gp, gn, clip, 1e-8)
        refb = torch.nn.functional.binary_cross_entropy_with_logits(sc, t.cpu() * 0.8 + 0.1, reduction="sum")
        (ref + refb).backward()
        assert _close(asl, ref) and _close(bce, refb), (gp, gn, clip)
        assert _close(got, sc.grad, 2e-4), (gp, gn, clip)
