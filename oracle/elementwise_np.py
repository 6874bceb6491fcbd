"""numpy restatement, with ANALYTIC gradients, of the element-wise part of the MVAE
step (test infrastructure; not on the product path).

These are the exact formulas the fused CUDA kernels implement; tests check them
against torch autograd of ``oracle/mvae_oracle.py`` (which in turn is pinned to the
reference by ``tests/golden``), and the CUDA kernels against them.

Reference lines restated:
  PoE            mnist/model.py:156-163 (variant A) ; celeba/model.py:200-207 (variant B)
  prior expert   mnist/model.py:172-185 (mu=0, logvar=0, always expert #0)
  reparametrize  mnist/model.py:29-35
  KL             mnist/train.py:56
  BCE logits     mnist/train.py:62-74
  CE             mnist/train.py:77-94
  mean over B    mnist/train.py:57-58
"""
from __future__ import annotations

import numpy as np


def swish(x):
    return x / (1.0 + np.exp(-x))


def swish_grad(x):
    s = 1.0 / (1.0 + np.exp(-x))
    return s * (1.0 + x * (1.0 - s))


def poe_multipass_fwd(mu_e, lv_e, masks, noise, beta, variant="A", eps=1e-8):
    """E encoder experts [E,B,L] (prior implicit), P passes given by bit masks over the E
    experts.  Returns per pass mu, logvar, z and the per-pass KL contribution
    beta * mean_b KL_b.  noise: [P,B,L] or None (eval: z = mu)."""
    dt = mu_e.dtype.type
    c1 = dt(2 if variant == "A" else 1)
    E, B, L = mu_e.shape
    P = len(masks)
    T_e = dt(1) / ((np.exp(lv_e) + dt(eps)) + (dt(eps) if variant == "A" else dt(0)))
    T0 = dt(1) / ((dt(1) + dt(eps)) + (dt(eps) if variant == "A" else dt(0)))
    mu = np.zeros((P, B, L), mu_e.dtype); lv = np.zeros_like(mu); z = np.zeros_like(mu)
    kl = np.zeros(P, np.float64)
    for p, m in enumerate(masks):
        S = np.full((B, L), T0, mu_e.dtype); N = np.zeros((B, L), mu_e.dtype)
        for e in range(E):
            if (m >> e) & 1:
                S = S + T_e[e]; N = N + mu_e[e] * T_e[e]
        mu[p] = N / S
        pv = dt(1) / S
        lv[p] = np.log(pv + (dt(eps) if variant == "A" else dt(0)))
        z[p] = mu[p] if noise is None else noise[p] * np.exp(dt(0.5) * lv[p]) + mu[p]
        klrow = -0.5 * np.sum(1 + lv[p].astype(np.float64) - mu[p].astype(np.float64) ** 2
                              - np.exp(lv[p].astype(np.float64)), axis=1)
        kl[p] = beta * klrow.mean()
    return mu, lv, z, kl


def poe_multipass_bwd(mu_e, lv_e, masks, noise, beta, dz, variant="A", eps=1e-8):
    """Gradient of  sum_p [ <dz_p, z_p> + beta*mean_b KL_p ]  w.r.t. the expert outputs.
    dz: [P,B,L] upstream gradient of the decoders w.r.t. z_p."""
    dt = mu_e.dtype.type
    E, B, L = mu_e.shape
    e2 = dt(eps) if variant == "A" else dt(0)
    ex = np.exp(lv_e)
    T_e = dt(1) / ((ex + dt(eps)) + e2)
    T0 = dt(1) / ((dt(1) + dt(eps)) + e2)
    dmu_e = np.zeros_like(mu_e); dlv_e = np.zeros_like(lv_e)
    for p, m in enumerate(masks):
        S = np.full((B, L), T0, mu_e.dtype); N = np.zeros((B, L), mu_e.dtype)
        for e in range(E):
            if (m >> e) & 1:
                S = S + T_e[e]; N = N + mu_e[e] * T_e[e]
        mu = N / S; pv = dt(1) / S; lv = np.log(pv + e2)
        g_mu = dz[p] + dt(beta / B) * mu
        g_lv = dt(beta / B) * dt(0.5) * (np.exp(lv) - dt(1))
        if noise is not None:
            g_lv = g_lv + dz[p] * noise[p] * dt(0.5) * np.exp(dt(0.5) * lv)
        # d lv / d S = -(pv^2)/(pv + e2)
        dS_from_lv = -(pv * pv) / (pv + e2)
        for e in range(E):
            if (m >> e) & 1:
                dmu_e[e] += g_mu * T_e[e] / S
                dT = g_mu * (mu_e[e] - mu) / S + g_lv * dS_from_lv
                dlv_e[e] += dT * (-(T_e[e] * T_e[e]) * ex[e])
    return dmu_e, dlv_e


def bce_logits_fwd_bwd(x, t, scale):
    """loss = scale * sum(max(x,0) - x t + log1p(exp(-|x|))) ; dx = scale*(sigmoid(x) - t).
    (scale = lambda / B).  loss accumulated in float64."""
    l = np.maximum(x, 0) - x * t + np.log(1 + np.exp(-np.abs(x)))
    dx = (1.0 / (1.0 + np.exp(-x.astype(np.float64))) - t).astype(x.dtype) * x.dtype.type(scale)
    return scale * float(l.astype(np.float64).sum()), dx


def ce_fwd_bwd(x, target, scale, eps=1e-6):
    """loss = scale * sum_b -log_softmax(x+eps)[b,target_b] ; dx = scale*(softmax - onehot)."""
    xs = (x + x.dtype.type(eps)).astype(np.float64)
    mx = xs.max(axis=1, keepdims=True)
    lse = mx + np.log(np.exp(xs - mx).sum(axis=1, keepdims=True))
    logp = xs - lse
    B = x.shape[0]
    loss = -logp[np.arange(B), target].sum() * scale
    dx = np.exp(logp)
    dx[np.arange(B), target] -= 1.0
    return float(loss), (dx * scale).astype(x.dtype)
