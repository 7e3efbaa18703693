"""Weight container of one LENS layer (inference branch only).

Mirrors lens/src/blitnet.py:41-64 of the reference: in inference mode an `SNNLayer` is a
bias-free `nn.Linear` called `w` plus an (unused at inference) threshold parameter `thr`,
which is what gives the trained checkpoints their four state-dict keys
(`feature_layer.thr`, `feature_layer.w.weight`, `output_layer.thr`, `output_layer.w.weight`).
Training (STDP, lens/src/blitnet.py:65-253) is out of scope for this package.
"""
import torch
import torch.nn as nn


class SNNLayer(nn.Module):
    def __init__(self, dims=[0, 0], thr_range=[0, 0], fire_rate=[0, 0], ip_rate=0, stdp_rate=0,
                 const_inp=[0, 0], p=[1, 1], spk_force=False, device=None, inference=False, args=None):
        super().__init__()
        self.device = device
        if not inference:
            raise NotImplementedError(
                "lens_b200 implements the inference hot path only; train with the reference "
                "(lens/train_model.py) and load the resulting .pth here")
        self.w = nn.Linear(dims[0], dims[1], bias=False)
        self.w.to(device)
        self.thr = nn.Parameter(torch.zeros([1, dims[-1]], device=device).uniform_(thr_range[0],
                                                                                   thr_range[1]))
