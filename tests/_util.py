"""Shared helpers for the tests (fixtures loading, state-dict rebuilding)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    out, sd, grads = {}, {}, {}
    for k in z.files:
        v = z[k]
        t = torch.from_numpy(v) if v.dtype.kind in "fiub" and v.ndim > 0 else v
        if k.startswith("sd::"):
            sd[k[4:]] = torch.as_tensor(v)
        elif k.startswith("grad::"):
            grads[k[6:]] = t
        else:
            out[k] = t if not (isinstance(t, np.ndarray) and t.ndim == 0) else t.item()
    out["sd"] = sd
    out["grads"] = grads
    return out
