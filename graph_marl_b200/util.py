"""Host helpers the reference's training script expects next to the rollout classes
(behaviour of src/util.py:8-111; own implementation).

Checkpoint layout kept so reference checkpoints load and ours load there:
    {"type": <model class name>, "state_dict": ..., "args": ..., ["netmon_state_dict": ...]}
"""
import random as _py_random

import numpy as _np
import torch as _torch

_CKPT_MODEL, _CKPT_NETMON = "state_dict", "netmon_state_dict"


def interpolate_model(a, b, a_weight, target):
    """target <- a_weight * a + (1 - a_weight) * b over every state_dict entry (soft target update,
    main.py:1019-1020 calls it with a=model, b=target=model_tar, a_weight=tau)."""
    theirs = b.state_dict()
    w_a, w_b = a_weight, 1 - a_weight
    target.load_state_dict({name: w_a * ours + w_b * theirs[name] for name, ours in a.state_dict().items()})


def get_state_dict(model, netmon, args):
    ckpt = dict(type=model.__class__.__name__, args=args)
    ckpt[_CKPT_MODEL] = model.state_dict()
    if netmon is not None:
        ckpt[_CKPT_NETMON] = netmon.state_dict()
    return ckpt


def load_state_dict(state_dict, model, netmon):
    expected, found = model.__class__.__name__, state_dict["type"]
    if expected != found:
        print(f"Warning: Loader expected {expected} but found {found}")
    has_netmon = _CKPT_NETMON in state_dict
    if has_netmon and netmon is None:
        raise ValueError("Model uses NetMon which has not been initialized.")
    if netmon is not None and not has_netmon:
        raise ValueError("NetMon state could not be found.")
    if has_netmon:
        netmon.load_state_dict(state_dict[_CKPT_NETMON])
    model.load_state_dict(state_dict[_CKPT_MODEL])


def set_attributes(obj, key_value_dict, verbose=False):
    """setattr for every pair; with `verbose`, one '> Updated/Added: k = v' line per changed attribute."""
    _missing = object()
    report = []
    for name, new in key_value_dict.items():
        old = getattr(obj, name, _missing)
        if old is _missing:
            report.append(f"> Added: {name} = {new}")
        elif old != new:
            report.append(f"> Updated: {name} = {new}")
        setattr(obj, name, new)
    if verbose and report:
        print("\n".join(report))


def filter_dict(dict, keys):
    return {k: dict[k] for k in keys}


def one_hot_list(i, max_indices):
    """List of `max_indices` zeros with a one at i; all zeros for a negative i."""
    return [int(j == i) for j in range(max_indices)] if i >= 0 else [0] * max_indices


def set_seed(seed):
    """Seeds torch, python `random` and the legacy numpy stream (the env and the policy draw from the latter)."""
    for seeder in (_torch.manual_seed, _py_random.seed, _np.random.seed):
        seeder(seed)


def dim_str_to_list(dims: str):
    """'512,256' -> [512, 256]; '' -> []."""
    return [int(tok) for tok in dims.split(",")] if dims else []
