"""Minimal stand-in for the parts of `sinabs` that LENS's inference path touches.

TEST INFRASTRUCTURE ONLY (used by make_golden.py in the build container, where
the real `sinabs` wheel is not installable: no network, no wheel in
/opt/wheelhouse).  It restates, in plain torch fp32 ops, the published
behaviour of sinabs>=2.0 for exactly the call sites the reference uses
(/root/reference/lens/run_model.py:34,44,151-156,236,238):

* ``sinabs.from_torch.from_model(model, input_shape, num_timesteps,
  add_spiking_output=True)``: appends a ReLU after the last weight layer,
  deep-copies the model and replaces every ``nn.ReLU`` by
  ``IAFSqueeze(spike_threshold=1.0, spike_fn=MultiSpike,
  reset_fn=MembraneSubtract(), min_v_mem=-1.0, num_timesteps=T)``.
* ``IAFSqueeze.forward``: [B'*T, ...] -> [B', T, ...]; per step
      v   = 1.0 * v + x_t                       (IAF: alpha_mem = 1, no synapse)
      s_t = (v > 0) * trunc(v / thr)            (MultiSpike)
      v   = v - s_t * thr                       (MembraneSubtract)
      v   = relu(v - v_min) + v_min             (min_v_mem clip, two fp32 roundings)
  state zero-initialised on first use / shape change, otherwise carried over
  between calls (the reference never calls reset_states()).
* ``sinabs.layers.FlattenTime``: [B, T, ...] -> [B*T, ...].

Because the real package is absent this part of the golden data is
"sinabs-restated" ("parity pinning").
"""
import copy
import sys
import types

import torch
import torch.nn as nn


class IAFSqueeze(nn.Module):
    def __init__(self, spike_threshold=1.0, min_v_mem=-1.0, num_timesteps=None):
        super().__init__()
        self.spike_threshold = torch.tensor(float(spike_threshold))
        self.min_v_mem = torch.tensor(float(min_v_mem))
        self.num_timesteps = num_timesteps
        self.v_mem = None
        self.record = None  # optional list collecting per-call output spikes

    def reset_states(self):
        if self.v_mem is not None:
            self.v_mem = torch.zeros_like(self.v_mem)

    def forward(self, x):
        T = self.num_timesteps
        n = x.shape[0]
        b = n // T
        x = x.reshape(b, T, *x.shape[1:])
        if self.v_mem is None or self.v_mem.shape != x[:, 0].shape:
            self.v_mem = torch.zeros_like(x[:, 0])
        v = self.v_mem
        alpha = torch.tensor(1.0)
        thr = self.spike_threshold
        out = []
        for t in range(T):
            v = alpha * v + x[:, t]
            s = (v > 0) * torch.div(v, thr, rounding_mode="trunc").float()
            v = v - s * thr
            v = torch.nn.functional.relu(v - self.min_v_mem) + self.min_v_mem
            out.append(s)
        self.v_mem = v
        y = torch.stack(out, 1).reshape(n, *x.shape[2:])
        if self.record is not None:
            self.record.append(y.clone())
        return y


class FlattenTime(nn.Flatten):
    def __init__(self):
        super().__init__(start_dim=0, end_dim=1)


class Network(nn.Module):
    def __init__(self, analog_model, spiking_model):
        super().__init__()
        self.analog_model = analog_model
        self.spiking_model = spiking_model

    def forward(self, x):
        return self.spiking_model(x)

    def reset_states(self):
        for m in self.spiking_model.modules():
            if isinstance(m, IAFSqueeze):
                m.reset_states()


def from_model(model, input_shape=None, spike_threshold=1.0, min_v_mem=-1.0,
               num_timesteps=None, add_spiking_output=False, **kw):
    layers = list(model.children())
    if add_spiking_output:
        layers = layers + [nn.ReLU()]
    spiking = copy.deepcopy(nn.Sequential(*layers))
    for i, m in enumerate(spiking):
        if isinstance(m, nn.ReLU):
            spiking[i] = IAFSqueeze(spike_threshold, min_v_mem, num_timesteps)
    net = Network(model, spiking)
    if input_shape is not None and num_timesteps is not None:
        with torch.no_grad():  # sinabs' constructor does one dummy pass, then resets
            dev = next(model.parameters()).device
            net(torch.zeros(num_timesteps, *input_shape, device=dev))
        net.reset_states()
    return net


def install():
    """Register fake `sinabs*` modules so that `import lens.run_model` succeeds."""
    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m
    s = mod("sinabs")
    sl = mod("sinabs.layers")
    sl.FlattenTime = FlattenTime
    sl.IAFSqueeze = IAFSqueeze
    s.layers = sl
    ft = mod("sinabs.from_torch")
    ft.from_model = from_model
    s.from_torch = ft
    be = mod("sinabs.backend")
    dc = mod("sinabs.backend.dynapcnn")
    dc.DynapcnnNetwork = object
    cf = mod("sinabs.backend.dynapcnn.chip_factory")
    cf.ChipFactory = object
    s.backend = be
    be.dynapcnn = dc
    dc.chip_factory = cf
