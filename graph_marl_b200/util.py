"""Small helpers with the reference's semantics (src/util.py:8-111)."""
import random

import numpy as np
import torch
import torch.nn as nn


def interpolate_model(a: nn.Module, b: nn.Module, a_weight: float, target: nn.Module):
    a_dict, b_dict = a.state_dict(), b.state_dict()
    for key in a_dict:
        a_dict[key] = a_weight * a_dict[key] + (1 - a_weight) * b_dict[key]
    target.load_state_dict(a_dict)


def get_state_dict(model, netmon, args):
    sd = {"type": type(model).__name__, "state_dict": model.state_dict(), "args": args}
    if netmon is not None:
        sd["netmon_state_dict"] = netmon.state_dict()
    return sd


def load_state_dict(state_dict, model, netmon):
    if state_dict["type"] != type(model).__name__:
        print(f"Warning: Loader expected {type(model).__name__} but found {state_dict['type']}")
    if "netmon_state_dict" in state_dict:
        if netmon is None:
            raise ValueError("Model uses NetMon which has not been initialized.")
        netmon.load_state_dict(state_dict["netmon_state_dict"])
    elif netmon is not None:
        raise ValueError("NetMon state could not be found.")
    model.load_state_dict(state_dict["state_dict"])


def set_attributes(obj, key_value_dict, verbose=False):
    for key, value in key_value_dict.items():
        if verbose and (not hasattr(obj, key) or getattr(obj, key) != value):
            print(f"> {'Updated' if hasattr(obj, key) else 'Added'}: {key} = {value}")
        setattr(obj, key, value)


def filter_dict(dict, keys):
    return {key: dict[key] for key in keys}


def one_hot_list(i, max_indices):
    a = [0] * max_indices
    if i >= 0:
        a[i] = 1
    return a


def set_seed(seed):
    torch.manual_seed(seed)
    random.seed(seed)
    np.random.seed(seed)


def dim_str_to_list(dims: str):
    return [] if len(dims) == 0 else [int(item) for item in dims.split(",")]
