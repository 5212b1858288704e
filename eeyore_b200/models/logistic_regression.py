"""Bayesian logistic (softmax) regression; mirror of eeyore/models/logistic_regression.py:8-37.
It is the single-layer special case of the MLP path and runs on the runtime-shape kernels (csrc/generic.cuh)."""
import torch

from .mlp import MLP
from .mlp import Hyperparameters as _MLPHyperparameters


class Hyperparameters:
    def __init__(self, input_size=1, output_size=1, bias=True, activation=torch.sigmoid):
        self.input_size, self.output_size, self.bias, self.activation = input_size, output_size, bias, activation


class _OneLayer(_MLPHyperparameters):
    def __init__(self, hp):
        self.dims, self.bias, self.activations = [hp.input_size, hp.output_size], [hp.bias], [hp.activation]


class LogisticRegression(MLP):
    def __init__(self, loss, temperature=None, prior=None, hparams=None, savefile=None, dtype=torch.float64, device=None):
        self.lr_hp = hparams if hparams is not None else Hyperparameters()
        super().__init__(loss, temperature=temperature, prior=prior, hparams=_OneLayer(self.lr_hp), savefile=savefile,
                         dtype=dtype, device=device)

    def __repr__(self):
        return f"LogisticRegression(input_size={self.lr_hp.input_size}, output_size={self.lr_hp.output_size}, dtype={self.dtype})"
