import torch
import torch.nn as nn

DEV = torch.device('cuda:0')  # utils/modelutils.py:4 of the reference


def find_layers(module, layers=(nn.Conv2d, nn.Linear), name=''):
    """{dotted name: module} for every sub-module whose exact type is in `layers` (utils/modelutils.py:7-13)."""
    found = {}
    stack = [(name, module)]
    while stack:
        prefix, mod = stack.pop()
        if type(mod) in tuple(layers):
            found[prefix] = mod
            continue
        for child_name, child in mod.named_children():
            stack.append((f'{prefix}.{child_name}' if prefix else child_name, child))
    return found
