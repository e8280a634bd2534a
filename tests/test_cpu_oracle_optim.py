"""Pins the oracle's restated keras.optimizers update rules (oracle/optim.py, [DEP]) against the independent implementations of the same
published algorithms in torch.optim, over 25 steps of a fixed gradient sequence:

  * SGD with momentum / Nesterov momentum: Keras keeps v = mu*v - lr*g and adds it, torch keeps buf = mu*buf + g and subtracts
    lr*buf -- the same trajectory for a constant learning rate;
  * RMSprop: identical form in both (a = rho*a + (1-rho)*g^2; p -= lr*g / (sqrt(a) + eps), eps outside the root);
  * Nadam: torch.optim.NAdam implements Dozat's formulation with the Keras momentum schedule mu_t = b1*(1 - 0.5*0.96^(t*decay));
  * Adam: the two differ ONLY in where epsilon sits (Keras: lr_t*m / (sqrt(v) + eps) with lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    torch: eps added after the bias correction of v), i.e. Keras-Adam(eps) at step t == torch-Adam(eps / sqrt(1-b2^t)); the
    test checks exactly that identity step by step, and plain equality for eps -> 0.
The engine's optimizer kernels are then held to the oracle on the GPU (tests/test_gpu_ops.py::test_optimizers)."""
import math

import pytest
import torch

from oracle import optim as OO


def _setup(seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = {"a/kernel": (3, 3, 8, 16), "a/bias": (16,), "bn/gamma": (16,)}
    p0 = {k: torch.randn(s, generator=g) for k, s in shapes.items()}
    grads = [{k: torch.randn(s, generator=g) * (0.5 + 0.1 * t) for k, s in shapes.items()} for t in range(25)]
    return p0, grads


def _run_torch(opt_fn, p0, grads):
    ps = {k: v.clone().requires_grad_(True) for k, v in p0.items()}
    opt = opt_fn(list(ps.values()))
    for g in grads:
        for k, p in ps.items():
            p.grad = g[k].clone()
        opt.step()
    return {k: p.detach() for k, p in ps.items()}


def _run_oracle(cls, p0, grads, **kw):
    ps = {k: v.clone() for k, v in p0.items()}
    opt = cls(ps, **kw)
    for g in grads:
        opt.step({k: v.clone() for k, v in g.items()})
    return ps


def _close(a, b, tol):
    for k in a:
        d = float((a[k] - b[k]).abs().max())
        assert d <= tol * (1.0 + float(b[k].abs().max())), (k, d)


@pytest.mark.parametrize("nesterov", [False, True])
def test_sgd_momentum_is_torch_sgd(nesterov):
    p0, grads = _setup(1)
    want = _run_torch(lambda ps: torch.optim.SGD(ps, lr=0.01, momentum=0.9, nesterov=nesterov), p0, grads)
    got = _run_oracle(OO.SGD, p0, grads, lr=0.01, momentum=0.9, nesterov=nesterov)
    _close(got, want, 2e-6)


def test_rmsprop_is_torch_rmsprop():
    p0, grads = _setup(2)
    want = _run_torch(lambda ps: torch.optim.RMSprop(ps, lr=1e-3, alpha=0.9, eps=1e-7), p0, grads)
    got = _run_oracle(OO.RMSprop, p0, grads, lr=1e-3, rho=0.9, epsilon=1e-7)
    _close(got, want, 2e-6)


def test_nadam_is_torch_nadam():
    p0, grads = _setup(3)
    want = _run_torch(lambda ps: torch.optim.NAdam(ps, lr=0.002, betas=(0.9, 0.999), eps=1e-7, momentum_decay=0.004), p0, grads)
    got = _run_oracle(OO.Nadam, p0, grads, lr=0.002, beta_1=0.9, beta_2=0.999, epsilon=1e-7, schedule_decay=0.004)
    _close(got, want, 2e-6)


def test_adam_is_torch_adam_up_to_the_epsilon_placement():
    p0, grads = _setup(4)
    # eps -> 0: the same algorithm
    want = _run_torch(lambda ps: torch.optim.Adam(ps, lr=1e-3, betas=(0.9, 0.999), eps=1e-30), p0, grads)
    got = _run_oracle(OO.Adam, p0, grads, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-30)
    _close(got, want, 2e-6)
    # eps = 1e-7 (K.epsilon()): Keras' step t equals torch's step with eps / sqrt(1 - b2^t).  A large epsilon relative to sqrt(v)
    # makes the placement visible: gradients of order 1e-6
    small = [{k: v * 1e-6 for k, v in g.items()} for g in grads]
    ps = {k: v.clone().requires_grad_(True) for k, v in p0.items()}
    topt = torch.optim.Adam(list(ps.values()), lr=1e-3, betas=(0.9, 0.999), eps=1e-7)
    oparams = {k: v.clone() for k, v in p0.items()}
    oopt = OO.Adam(oparams, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7)
    for t, g in enumerate(small, start=1):
        for grp in topt.param_groups:
            grp["eps"] = 1e-7 / math.sqrt(1.0 - 0.999 ** t)
        for k, p in ps.items():
            p.grad = g[k].clone()
        topt.step()
        oopt.step({k: v.clone() for k, v in g.items()})
    _close(oparams, {k: p.detach() for k, p in ps.items()}, 2e-6)
    # and with torch's own (fixed) epsilon the trajectories DO separate on such gradients -- the placement is not a no-op
    fixed = _run_torch(lambda q: torch.optim.Adam(q, lr=1e-3, betas=(0.9, 0.999), eps=1e-7), p0, small)
    assert max(float((fixed[k] - oparams[k]).abs().max()) for k in fixed) > 1e-4


def test_loss_formulas_against_independent_implementations():
    """binary / categorical cross-entropy against torch's own losses on Keras-clipped probabilities (clip 1e-7), and the focal-loss
    FORMULA -alpha*(1-p)^gamma*t*log p - (1-alpha)*p^gamma*(1-t)*log(1-p) against torchvision.ops.sigmoid_focal_loss (what stays a
    named flag is only the reduction the absent musket_core applies, oracle/losses.py)"""
    import torch.nn.functional as F
    from oracle import losses as OL
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(3, 16, 20, 1, generator=g) * 3
    p = torch.sigmoid(logits)
    t = (torch.rand(3, 16, 20, 1, generator=g) > 0.6).float()
    pc = p.clamp(1e-7, 1 - 1e-7)
    assert abs(float(OL.binary_crossentropy(t, p)) - float(F.binary_cross_entropy(pc, t))) < 1e-6
    z = torch.randn(3, 16, 20, 4, generator=g)
    pm = torch.softmax(z, dim=-1)
    onehot = F.one_hot(torch.randint(0, 4, (3, 16, 20), generator=g), 4).float()
    want = F.nll_loss(torch.log(pm.clamp(1e-7, 1.0)).reshape(-1, 4), onehot.argmax(-1).reshape(-1))
    assert abs(float(OL.categorical_crossentropy(onehot, pm)) - float(want)) < 1e-6
    tvo = pytest.importorskip("torchvision.ops")
    for alpha, gamma in ((0.75, 2.0), (0.25, 2.0), (0.5, 1.0)):
        for red in ("mean", "sum"):
            want = tvo.sigmoid_focal_loss(logits, t, alpha=alpha, gamma=gamma, reduction=red)
            got = OL.focal_loss(t, p, gamma=gamma, alpha=alpha, reduction=red)
            assert abs(float(got) - float(want)) <= 2e-5 * (1.0 + abs(float(want))), (alpha, gamma, red, float(got), float(want))


def test_lovasz_hinge_is_the_lovasz_extension_of_the_jaccard_loss():
    """Berman et al. 2018 (eq. 8-9) by its DEFINITION, with Python sets: sort the hinge errors m decreasingly (permutation pi); the loss
    is sum_i m_pi(i) * [Delta_J({pi_1..pi_i}) - Delta_J({pi_1..pi_(i-1)})] with Delta_J(M) = 1 - |P \\ M| / |P u (M n N)| the Jaccard
    loss when exactly the pixels in M are mispredicted (P / N: ground-truth positives / negatives).  The oracle's cumulative-sum form
    (`_lovasz_grad`) must give the same number; act='relu' is the paper's hinge, act='elu' only changes m -> elu(m) + 1."""
    from oracle import losses as OL
    g = torch.Generator().manual_seed(9)
    for trial in range(6):
        n = 14
        logits = torch.randn(n, generator=g) * 2
        labels = (torch.rand(n, generator=g) > (0.3 + 0.1 * trial)).float()
        if trial == 5:
            labels.zero_()                                   # no positive pixel at all
        signs = 2 * labels - 1
        errors = (1 - logits * signs)
        order = sorted(range(n), key=lambda i: -float(errors[i]))
        P = {i for i in range(n) if labels[i] == 1}
        N = set(range(n)) - P

        def delta(M):
            union = len(P | (M & N))
            return 1.0 - (len(P - M) / union) if union else 0.0   # (empty union only for M = {} and no positives)
        want, M, prev = 0.0, set(), 0.0
        for i in order:
            M = M | {i}
            d = delta(M)
            want += max(float(errors[i]), 0.0) * (d - prev)
            prev = d
        got = float(OL.lovasz_hinge_flat(logits, labels, act="relu"))
        assert abs(got - want) < 1e-5 * (1 + abs(want)), (trial, got, want)
        e = torch.tensor([float(errors[i]) for i in order])
        want_elu, M, prev = 0.0, set(), 0.0
        for k, i in enumerate(order):
            M = M | {i}
            d = delta(M)
            want_elu += float(torch.nn.functional.elu(e[k]) + 1.0) * (d - prev)
            prev = d
        assert abs(float(OL.lovasz_hinge_flat(logits, labels, act="elu")) - want_elu) < 1e-5 * (1 + abs(want_elu))
