"""Shared helpers for the test-suite (fixtures come from oracle/ + tests/golden/)."""
import os

import numpy as np
import torch

from oracle import unet_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def fixture_sd(out_ch=10):
    """Fixture F1: seed-per-tensor synthetic weights + the BN running stats frozen in tests/golden."""
    sd = O.synth_state_dict(O.mbv2unet_param_shapes(out_ch), seed=0)
    bn = gold("mbv2unet_bnstats.npz")
    for k in bn.files:
        sd[k] = torch.from_numpy(bn[k]).clone()
    return sd


def expand_aliases(sd):
    """Add the downK.N.* spellings (unet.py:15-19) so load_state_dict(strict=True) accepts it."""
    out = dict(sd)
    for k, v in sd.items():
        if k.startswith("backbone.features."):
            n = int(k.split(".")[2])
            d = 1 if n < 2 else 2 if n < 4 else 3 if n < 7 else 4 if n < 11 else 5
            out[f"down{d}." + k[len("backbone.features."):]] = v
    return out


def rel_err(a, b):
    """max |a-b| / max |b|  -- the 'relative' of BASELINE's tolerances."""
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def f2_sd():
    """Fixture F2: the network the LIVE reference trained for 600 Adam steps on the road-scene generator
    (oracle/make_golden_f2.py), reconstructed exactly from tests/golden/f2_weights.npz."""
    return O.f2_state_dict(gold("f2_weights.npz"))


def f2_meta():
    import json
    return json.loads(str(gold("f2_eval.npz")["meta"]))
