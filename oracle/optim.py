"""keras.optimizers (2.2.4) restated [DEP]; reference selects them via `optimizer:`/`lr:`/`clipnorm:`/`clipvalue:`
(schemas/segmentation.raml:77-89).  TEST INFRASTRUCTURE.  SURVEY.md 8 a-9.  The Keras source is absent, so the rules are restated
from its published formulation; they are pinned against torch.optim's independent implementations of the same algorithms
(tests/test_cpu_oracle_optim.py: SGD / Nesterov, RMSprop, Nadam equal; Adam equal up to Keras' documented epsilon placement).
"""
from __future__ import annotations

import math
from typing import Dict

import torch


def _clip(grads: Dict[str, torch.Tensor], clipnorm=None, clipvalue=None):
    if clipnorm is not None and clipnorm > 0:
        norm = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads.values()))
        if norm > clipnorm:
            grads = {k: g * (clipnorm / norm) for k, g in grads.items()}
    if clipvalue is not None and clipvalue > 0:
        grads = {k: g.clamp(-clipvalue, clipvalue) for k, g in grads.items()}
    return grads


class Adam:
    """Keras Adam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps)  (eps OUTSIDE the bias
    correction, eps=1e-7 = K.epsilon(); differs from torch.optim.Adam)."""

    def __init__(self, params: Dict[str, torch.Tensor], lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7,
                 clipnorm=None, clipvalue=None):
        self.params, self.lr, self.b1, self.b2, self.eps = params, lr, beta_1, beta_2, epsilon
        self.clipnorm, self.clipvalue = clipnorm, clipvalue
        self.t = 0
        self.m = {k: torch.zeros_like(p) for k, p in params.items()}
        self.v = {k: torch.zeros_like(p) for k, p in params.items()}

    @torch.no_grad()
    def step(self, grads: Dict[str, torch.Tensor]):
        grads = _clip(grads, self.clipnorm, self.clipvalue)
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        for k, p in self.params.items():
            g = grads.get(k)
            if g is None:
                continue
            self.m[k].mul_(self.b1).add_(g, alpha=1 - self.b1)
            self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            p.sub_(lr_t * self.m[k] / (self.v[k].sqrt() + self.eps))


class SGD:
    """Keras SGD: v = mu*v - lr*g; p += v (nesterov: p += mu*v - lr*g)."""

    def __init__(self, params, lr=0.01, momentum=0.0, nesterov=False, clipnorm=None, clipvalue=None):
        self.params, self.lr, self.mu, self.nesterov = params, lr, momentum, nesterov
        self.clipnorm, self.clipvalue = clipnorm, clipvalue
        self.v = {k: torch.zeros_like(p) for k, p in params.items()}

    @torch.no_grad()
    def step(self, grads):
        grads = _clip(grads, self.clipnorm, self.clipvalue)
        for k, p in self.params.items():
            g = grads.get(k)
            if g is None:
                continue
            self.v[k].mul_(self.mu).sub_(g, alpha=self.lr)
            if self.nesterov:
                p.add_(self.mu * self.v[k] - self.lr * g)
            else:
                p.add_(self.v[k])


class RMSprop:
    """Keras RMSprop: a = rho*a+(1-rho)g^2; p -= lr*g/(sqrt(a)+eps); rho=0.9, eps=1e-7."""

    def __init__(self, params, lr=1e-3, rho=0.9, epsilon=1e-7, clipnorm=None, clipvalue=None):
        self.params, self.lr, self.rho, self.eps = params, lr, rho, epsilon
        self.clipnorm, self.clipvalue = clipnorm, clipvalue
        self.a = {k: torch.zeros_like(p) for k, p in params.items()}

    @torch.no_grad()
    def step(self, grads):
        grads = _clip(grads, self.clipnorm, self.clipvalue)
        for k, p in self.params.items():
            g = grads.get(k)
            if g is None:
                continue
            self.a[k].mul_(self.rho).addcmul_(g, g, value=1 - self.rho)
            p.sub_(self.lr * g / (self.a[k].sqrt() + self.eps))


class Nadam:
    """keras.optimizers.Nadam 2.2.4 [DEP] (segmentation.raml:77-89 optimizer enum): Nesterov Adam with the momentum
    schedule mu_t = b1*(1 - 0.5*0.96^(t*schedule_decay)); lr 0.002, eps 1e-7, schedule_decay 0.004."""

    def __init__(self, params, lr=0.002, beta_1=0.9, beta_2=0.999, epsilon=1e-7, schedule_decay=0.004, clipnorm=None,
                 clipvalue=None):
        self.params, self.lr, self.b1, self.b2, self.eps, self.sd = params, lr, beta_1, beta_2, epsilon, schedule_decay
        self.clipnorm, self.clipvalue = clipnorm, clipvalue
        self.t, self.m_schedule = 0, 1.0
        self.m = {k: torch.zeros_like(p) for k, p in params.items()}
        self.v = {k: torch.zeros_like(p) for k, p in params.items()}

    @torch.no_grad()
    def step(self, grads):
        grads = _clip(grads, self.clipnorm, self.clipvalue)
        self.t += 1
        t = self.t
        mu_t = self.b1 * (1.0 - 0.5 * 0.96 ** (t * self.sd))
        mu_t1 = self.b1 * (1.0 - 0.5 * 0.96 ** ((t + 1) * self.sd))
        ms_new = self.m_schedule * mu_t
        ms_next = ms_new * mu_t1
        self.m_schedule = ms_new
        for k, p in self.params.items():
            g = grads.get(k)
            if g is None:
                continue
            g_prime = g / (1.0 - ms_new)
            self.m[k].mul_(self.b1).add_(g, alpha=1 - self.b1)
            m_prime = self.m[k] / (1.0 - ms_next)
            self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            v_prime = self.v[k] / (1.0 - self.b2 ** t)
            m_bar = (1.0 - mu_t) * g_prime + mu_t1 * m_prime
            p.sub_(self.lr * m_bar / (v_prime.sqrt() + self.eps))


def make(name: str, params, lr=None, **kw):
    name = (name or "Adam").lower()
    if name == "adam":
        return Adam(params, lr=lr if lr is not None else 1e-3, **kw)
    if name == "sgd":
        return SGD(params, lr=lr if lr is not None else 0.01, **kw)
    if name == "rmsprop":
        return RMSprop(params, lr=lr if lr is not None else 1e-3, **kw)
    if name == "nadam":
        return Nadam(params, lr=lr if lr is not None else 0.002, **kw)
    raise ValueError("unknown optimizer " + name)
