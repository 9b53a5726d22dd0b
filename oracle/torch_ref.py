"""torch-CPU restatement of the reference training step -- TEST INFRASTRUCTURE ONLY.

Purpose (a) an autograd-based cross-check of the closed-form gradients in
``oracle/dca_oracle.py`` (autograd plays the role TF autodiff plays in the reference);
(b) the CPU baseline timed by ``bench.py`` ("torch-CPU restatement of the reference
path (TensorFlow unavailable in image)").

The op sequence mirrors what Keras executes for
  dca/network.py:92-141 (Dense -> BatchNormalization(center, no scale) -> relu),
  dca/network.py:366-393 / 496-522 / 293-316 (heads),  dca/layers.py:85 (mean * sf),
  dca/loss.py:72-156 (NB / ZINB NLL, every epsilon kept),
  dca/train.py:54-57 (RMSprop with clipvalue).
Never imported by the product package.
"""
from __future__ import annotations

import math
from typing import Dict, Sequence

import torch

from .dca_oracle import KERAS_DEFAULTS, layer_names, head_names

EPS = 1e-10


def nb_elem(y, mu, theta):
    theta = torch.clamp(theta, max=1e6)                                   # dca/loss.py:85
    t1 = torch.lgamma(theta + EPS) + torch.lgamma(y + 1.0) - torch.lgamma(y + theta + EPS)   # :87
    t2 = (theta + y) * torch.log(1.0 + (mu / (theta + EPS))) \
        + (y * (torch.log(theta + EPS) - torch.log(mu + EPS)))            # :88
    final = t1 + t2
    return torch.where(torch.isnan(final), torch.full_like(final, float("inf")), final)


def zinb_elem(y, mu, theta, pi, ridge=0.0):
    nb_case = nb_elem(y, mu, theta) - torch.log(1.0 - pi + EPS)           # :130
    th = torch.clamp(theta, max=1e6)
    zero_nb = torch.pow(th / (th + mu + EPS), th)                         # :136
    zero_case = -torch.log(pi + ((1.0 - pi) * zero_nb) + EPS)             # :137
    res = torch.where(y < 1e-8, zero_case, nb_case)                       # :138
    return res + ridge * pi * pi                                          # :139-140


class TorchRefNet:
    """Same parameter names / layouts as oracle.dca_oracle.OracleNet."""

    def __init__(self, params: Dict[str, "torch.Tensor"], hidden: Sequence[int], ae_type: str,
                 batchnorm=True, ridge=0.0, dtype=torch.float32):
        self.hidden = tuple(hidden); self.ae_type = ae_type; self.batchnorm = batchnorm
        self.ridge = ridge; self.dtype = dtype
        self.names = layer_names(len(self.hidden)); self.heads = head_names(ae_type)
        self.p = {k: torch.as_tensor(v).to(dtype).clone() for k, v in params.items()}
        self.train_keys = [k for k in self.p if k.endswith(("/kernel", "/bias", "/bn_beta", "/theta"))]
        for k in self.train_keys:
            self.p[k].requires_grad_(True)
        self.rms = {k: torch.zeros_like(self.p[k]) for k in self.train_keys}
        self.mom = KERAS_DEFAULTS["bn_momentum"]; self.bn_eps = KERAS_DEFAULTS["bn_eps"]

    def forward(self, X, sf, training=True):
        h = X
        stats = []
        for nm in self.names:
            a = h @ self.p[nm + "/kernel"] + self.p[nm + "/bias"]
            if self.batchnorm:
                if training:
                    mean = a.mean(0); var = a.var(0, unbiased=False)
                    stats.append((nm, mean.detach(), var.detach()))
                else:
                    mean = self.p[nm + "/bn_moving_mean"]; var = self.p[nm + "/bn_moving_var"]
                a = (a - mean) / torch.sqrt(var + self.bn_eps) + self.p[nm + "/bn_beta"]
            h = torch.relu(a)
        z = {nm: h @ self.p[nm + "/kernel"] + self.p[nm + "/bias"] for nm in self.heads}
        m = torch.clamp(torch.exp(z["mean"]), 1e-5, 1e6)
        mu = m * sf.reshape(-1, 1)
        if "dispersion" in z:
            theta = torch.clamp(torch.nn.functional.softplus(z["dispersion"]), 1e-4, 1e4)
        else:
            theta = torch.clamp(torch.exp(self.p["dispersion/theta"]), 1e-3, 1e4).reshape(1, -1)
        pi = torch.sigmoid(z["pi"]) if "pi" in z else None
        return mu, theta, pi, stats

    def loss(self, X, Y, sf, training=True):
        mu, theta, pi, stats = self.forward(X, sf, training)
        theta = theta.expand_as(mu)
        if pi is not None:
            el = zinb_elem(Y, mu, theta, pi, self.ridge)
        else:
            el = nb_elem(Y, mu, theta)
        return el.mean(), stats

    def loss_and_grads(self, X, Y, sf):
        for k in self.train_keys:
            self.p[k].grad = None
        loss, stats = self.loss(X, Y, sf, True)
        loss.backward()
        return float(loss.detach()), {k: self.p[k].grad.detach().clone() for k in self.train_keys}, stats

    @torch.no_grad()
    def _apply(self, grads, stats, lr, clip):
        rho = KERAS_DEFAULTS["rms_rho"]; eps = KERAS_DEFAULTS["rms_eps"]
        for k, g in grads.items():
            g = g.clamp(-clip, clip)
            self.rms[k].mul_(rho).addcmul_(g, g, value=1.0 - rho)
            self.p[k].sub_(lr * g / (self.rms[k].sqrt() + eps))
        for nm, mean, var in stats:
            self.p[nm + "/bn_moving_mean"].mul_(self.mom).add_((1 - self.mom) * mean)
            self.p[nm + "/bn_moving_var"].mul_(self.mom).add_((1 - self.mom) * var)

    def train_step(self, X, Y, sf, lr=KERAS_DEFAULTS["rms_lr"], clip=KERAS_DEFAULTS["clipvalue"]):
        loss, grads, stats = self.loss_and_grads(X, Y, sf)
        self._apply(grads, stats, lr, clip)
        return loss
