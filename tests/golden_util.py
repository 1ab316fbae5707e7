"""Helpers to read tests/golden/*.npz (produced by oracle/make_golden.py from the reference)."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(bytes(z["__meta__"]).decode())
    arrays = {k: z[k] for k in z.files if k != "__meta__"}
    return arrays, meta


def oracle_net(arrays, prefix, acts, requires_grad=False):
    W, b = [], []
    i = 0
    while f"{prefix}.W{i}" in arrays:
        W.append(torch.from_numpy(arrays[f"{prefix}.W{i}"].copy()).requires_grad_(requires_grad))
        b.append(torch.from_numpy(arrays[f"{prefix}.b{i}"].copy()).requires_grad_(requires_grad))
        i += 1
    return {"W": W, "b": b, "act": list(acts)}


def load_into_module(arrays, prefix, module):
    """Copy golden weights into a reagent_b200 model (module.fc.dnn[i][0] Linear views)."""
    fc = module.fc if hasattr(module, "fc") else module
    with torch.no_grad():
        for i, seq in enumerate(fc.dnn):
            seq[0].weight.copy_(torch.from_numpy(arrays[f"{prefix}.W{i}"]))
            seq[0].bias.copy_(torch.from_numpy(arrays[f"{prefix}.b{i}"]))


def batch_tensors(arrays, device="cpu"):
    return {k[len("batch."):]: torch.from_numpy(v.copy()).to(device)
            for k, v in arrays.items() if k.startswith("batch.")}


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
