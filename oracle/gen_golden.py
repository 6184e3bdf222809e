"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the authoring container (needs /root/reference or $GKG_REF):

    python oracle/gen_golden.py

The reference ships no golden vectors for this path, so the
fixtures are outputs of the reference's own classes (imported through
oracle/ref_shim.py) on seeded synthetic inputs.  They pin oracle/gkg_oracle.py and,
through it, the CUDA path.  Fixtures are kept small (a few hundred KB in total).
"""
from __future__ import annotations

import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def det_tensor(key, shape, dtype=torch.float32):
    """Deterministic parameter/buffer value keyed by its state-dict name.

    Shared by the generator and the tests (tests import it from here) so whole-model
    state dicts never have to be stored."""
    shape = tuple(int(s) for s in shape)
    if key.endswith("num_batches_tracked"):
        return torch.zeros(shape, dtype=torch.long)
    g = torch.Generator().manual_seed(zlib.crc32(key.encode()) & 0x7FFFFFFF)
    r = torch.randn(shape, generator=g, dtype=torch.float32)
    if key.endswith("running_var"):
        return (r.abs() * 0.25 + 0.75).to(dtype)
    if key.endswith("running_mean"):
        return (0.1 * r).to(dtype)
    if len(shape) >= 2:                       # conv / linear / embedding weights
        fan_in = int(np.prod(shape[1:]))
        if key == "pos_embed":
            return (0.02 * r).to(dtype)
        if key == "label_lt.weight":
            return r.to(dtype)
        return (r / fan_in ** 0.5).to(dtype)
    if key.endswith("weight"):                # norm scale
        return (1.0 + 0.1 * r).to(dtype)
    return (0.1 * r).to(dtype)                # biases


def fill_state_dict(module, keep=("relative_pos",)):
    sd = module.state_dict()
    new = {}
    for key, val in sd.items():
        if any(key.endswith(s) for s in keep):
            new[key] = val.clone()
        else:
            new[key] = det_tensor(key, val.shape, val.dtype)
    module.load_state_dict(new)
    return new


def npz(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    conv = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        conv[k] = v
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **conv)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KB)")


def sd_arrays(sd, prefix="sd::"):
    return {prefix + k: v for k, v in sd.items()}


def main():
    ref = ref_shim.load_reference()
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(1234)

    # ---- 1. kNN graph, separate keys, bias, dilation ---------------------------------
    x = torch.randn(4, 8, 64, 1, generator=g)
    y = torch.randn(4, 8, 16, 1, generator=g)
    rel = -(0.5 + 0.5 * torch.rand(1, 64, 16, generator=g))
    graph = ref.DenseDilatedKnnGraph(3, 2, False, 0.0)
    ei = graph(x, y, rel)
    xn = torch.nn.functional.normalize(x, p=2.0, dim=1).transpose(2, 1).squeeze(-1)
    yn = torch.nn.functional.normalize(y, p=2.0, dim=1).transpose(2, 1).squeeze(-1)
    dist = ref.xy_pairwise_distance(xn, yn) + rel
    npz("knn_xy", x=x, y=y, relative_pos=rel, k=3, dilation=2, edge_index=ei, dist=dist)

    # ---- 2. kNN graph, self keys ------------------------------------------------------
    x = torch.randn(4, 12, 49, 1, generator=g)
    rel = -(0.5 + 0.5 * torch.rand(1, 49, 49, generator=g))
    graph = ref.DenseDilatedKnnGraph(4, 3, False, 0.0)
    ei = graph(x, None, rel)
    ei_nobias = ref.DenseDilatedKnnGraph(4, 1, False, 0.0)(x)
    npz("knn_self", x=x, relative_pos=rel, k=4, dilation=3, edge_index=ei,
        edge_index_nobias_d1=ei_nobias)

    # ---- 3. MRConv2d (gather / max-relative / interleave / grouped FC + BN + GELU) ----
    conv = ref.MRConv2d(16, 32, "gelu", "batch", True).eval()
    sd = fill_state_dict(conv)
    x = torch.randn(4, 8, 64, 1, generator=g)           # B=2, G=2, D=8
    y = torch.randn(4, 8, 16, 1, generator=g)
    ei = ref.DenseDilatedKnnGraph(3, 1, False, 0.0)(x, y)
    out = conv(x, ei, y)
    out_self = conv(x, ref.DenseDilatedKnnGraph(3, 1, False, 0.0)(x), None)
    # pre-FC aggregate, recomputed with the reference's own gather
    x_i = ref.batched_index_select(x, ei[1])
    x_j = ref.batched_index_select(y, ei[0])
    npz("mrconv", x=x, y=y, edge_index=ei, out=out, out_self=out_self,
        x_j=x_j, maxrel=torch.max(x_j - x_i, -1, keepdim=True)[0], **sd_arrays(sd))

    # ---- 3b. the other GraphConv2d variants (torch_vertex.py:16-150), un-grouped as the reference builds them ----
    gv = torch.Generator().manual_seed(4321)          # own stream: the fixtures below keep their inputs
    xv = torch.randn(2, 16, 64, 1, generator=gv)
    yv = torch.randn(2, 16, 16, 1, generator=gv)
    ei_xy = ref.DenseDilatedKnnGraph(3, 1, False, 0.0)(xv, yv)
    ei_self = ref.DenseDilatedKnnGraph(4, 1, False, 0.0)(xv)
    for conv_name in ("edge", "sage", "gin", "gat"):
        m = ref.GraphConv2d(16, 32, conv_name, "gelu", "batch", True).eval()
        sdv = fill_state_dict(m)
        if conv_name == "gin":
            with torch.no_grad():
                m.gconv.eps.fill_(0.25)
            sdv = {k_: v_.clone() for k_, v_ in m.state_dict().items()}
        npz("gconv_" + conv_name, x=xv, y=yv, edge_index=ei_xy, edge_index_self=ei_self, out=m(xv, ei_xy, yv),
            out_self=m(xv, ei_self, None), **sd_arrays(sdv))

    # ---- 4./5. Grapher with key reduction (r=2) and without (r=1, dilation 2) --------
    for name, r, dil in (("grapher_r2", 2, 1), ("grapher_r1", 1, 2)):
        m = ref.Grapher(16, 3, dil, "mr", "gelu", "batch", True, False, 0.2, r, 64, 0.0,
                        True, True, 2).eval()
        sd = fill_state_dict(m)
        x = torch.randn(2, 16, 8, 8, generator=g)
        out = m(x)
        _, ei = m.graph_conv(m.fc1(x), m.relative_pos)
        npz(name, x=x, out=out, edge_index=ei, k=3, dilation=dil, r=r, **sd_arrays(sd))

    # non-grouped Grapher (DyGraphConv2d path)
    m = ref.Grapher(16, 3, 1, "mr", "gelu", "batch", True, False, 0.2, 2, 64, 0.0,
                    True, False, 2).eval()
    sd = fill_state_dict(m)
    x = torch.randn(2, 16, 8, 8, generator=g)
    npz("grapher_nogroup", x=x, out=m(x), k=3, dilation=1, r=2, **sd_arrays(sd))

    # ---- 6. GrapherLabel (label <-> patch group kNN head) -----------------------------
    for name, mg in (("grapher_label", True), ("grapher_label_nogroup", False)):
        m = ref.GrapherLabel(16, 3, 1, "mr", "gelu", "batch", True, False, 0.2, 1, 64, 0.0,
                             False, 5, mg, 2).eval()
        sd = fill_state_dict(m)
        lab = torch.randn(2, 5, 16, generator=g)
        feats = torch.randn(2, 16, 8, 8, generator=g)
        out, ei = m(lab, feats)
        npz(name, labels=lab, features=feats, out=out, edge_index=ei, k=3, **sd_arrays(sd))

    # ---- 7. relative position tables --------------------------------------------------
    tabs = {}
    for (c, n, r) in ((16, 64, 2), (16, 64, 1), (80, 144, 4), (48, 36, 1)):
        m = ref.Grapher(c, 3, 1, "mr", "gelu", "batch", True, False, 0.2, r, n, 0.0, True, True, 2)
        tabs[f"rel_c{c}_n{n}_r{r}"] = m.relative_pos.data
    npz("relative_pos", **tabs)

    # ---- 8. training-mode Grapher: outputs + input / parameter gradients --------------
    m = ref.Grapher(16, 3, 1, "mr", "gelu", "batch", True, False, 0.2, 2, 64, 0.0,
                    True, True, 2).train()
    sd = fill_state_dict(m)
    x = torch.randn(3, 16, 8, 8, generator=g, requires_grad=True)
    w = torch.randn(3, 16, 8, 8, generator=g)
    out = m(x)
    (out * w).sum().backward()
    grads = {"grad::" + k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    npz("grapher_train", x=x, w=w, out=out, grad_x=x.grad, k=3, dilation=1, r=2,
        **sd_arrays(sd), **grads)

    # ---- 9. whole backbone, GKGNet('s') at size 192, 7 classes ------------------------
    with ref_shim.cpu_cuda_noop():
        net = ref.GKGNet(choice="s", k=9, k_label_gcn=9, drop_path=0.0, n_classes=7,
                         size=192)
        net.eval()          # BaseBackbone.train() returns None (base_backbone.py:26-33)
        sd = fill_state_dict(net)
        img = det_tensor("img", (2, 3, 192, 192)) * (3 * 192 * 192) ** 0.5   # ~N(0,1); not stored
        with torch.no_grad():
            lab, gap, ei = net(img)
    keys = np.array([k for k in sd.keys()])
    shapes = np.array([",".join(str(s) for s in v.shape) for v in sd.values()])
    npz("gkgnet_s192", label_emb=lab, gap=gap, edge_index=ei, keys=keys, shapes=shapes,
        n_classes=7, size=192)


if __name__ == "__main__":
    main()
