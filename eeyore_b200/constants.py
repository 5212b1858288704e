"""Mirror of eeyore/constants/constants.py:7-18.

The reference's ``loss_functions`` values are python lambdas over torch ops; here they are tagged callables so the
MLP can select the native loss.  Calling one evaluates the same formula with torch ops on whatever device the
inputs live on (used only for API compatibility, never by the samplers).
"""
import numpy as np
import torch

from ._native import LOSS_BINARY, LOSS_MULTICLASS

torch_to_np_types = {torch.float32: np.float32, torch.float64: np.float64}


class NativeLoss:
    def __init__(self, name, loss_id):
        self.name, self.loss_id = name, loss_id

    def __call__(self, out, y):
        if self.loss_id == LOSS_BINARY:      # eeyore/stats/loss.py:1-11 with reduction='sum'
            return -(out.log() * y + (1 - out).log() * (1 - y)).sum()
        return torch.nn.functional.cross_entropy(out, torch.argmax(y, 1), reduction="sum")  # constants.py:17

    def __repr__(self):
        return f"NativeLoss({self.name})"


loss_functions = {
    "binary_classification": NativeLoss("binary_classification", LOSS_BINARY),
    "multiclass_classification": NativeLoss("multiclass_classification", LOSS_MULTICLASS),
}
