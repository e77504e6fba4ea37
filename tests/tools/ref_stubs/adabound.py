"""Stub of the `adabound` package the reference imports unconditionally (optimizers/__init__.py:4); tests only."""
import torch


class AdaBound(torch.optim.Adam):
    pass
