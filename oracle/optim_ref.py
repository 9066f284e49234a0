"""ORACLE (test infrastructure, never on the product path): numpy restatement of the step right after backward --
reference train.py:413-421:  scaler.unscale_(optimizer); clip_grad_norm_(params, 10.0); scaler.step(optimizer);
optimizer.zero_grad(); ema.update(model)  with optimizer = torch.optim.SGD(momentum, nesterov=True) over the three
parameter groups of utils/torch_utils.py:256-289 and ModelEMA.update of utils/torch_utils.py:359-368.  float32 arithmetic.

Only tests/ (and smoke / bench baseline legs) may import this.  Pinned: tests/golden/optim.npz holds the result of the real
reference objects (torch.optim.SGD, torch.nn.utils.clip_grad_norm_, the reference's ModelEMA through tests/golden/refshim.py)
on seeded tensors; tests/test_oracle_golden.py checks this file against it.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32


def sgd_ema_step(params, grads, moms, emas, groups, hyper, inv_scale=1.0, max_norm=10.0, ema_decay=0.9999, ema_tau=2000.0, ema_updates=0,
                 buffers=(), ema_buffers=()):
    """params/grads/moms/emas: lists of float32 arrays (updated copies are returned); groups[i] = index into `hyper`, a list of
    dicts(lr, momentum, weight_decay, nesterov).  Returns (params, moms, emas, ema_buffers, grad_norm, skipped)."""
    g = [F(inv_scale) * x.astype(F) for x in grads]                                    # scaler.unscale_
    finite = all(np.isfinite(x).all() for x in g)
    total = F(math.sqrt(sum(float((x.astype(np.float64) ** 2).sum()) for x in g)))     # clip_grad_norm_: norm of norms
    coef = F(1.0)
    if max_norm and max_norm > 0:
        coef = min(F(max_norm) / (total + F(1e-6)), F(1.0))
    p_out, m_out = [], []
    for p, gi, m, grp in zip(params, g, moms, groups):
        h = hyper[grp]
        if not finite:                                                                # GradScaler.step skips the whole update
            p_out.append(p.copy()); m_out.append(m.copy()); continue
        d = gi * coef
        if h["weight_decay"]:
            d = d + F(h["weight_decay"]) * p
        buf = F(h["momentum"]) * m + d
        d = d + F(h["momentum"]) * buf if h["nesterov"] else buf
        p_out.append((p - F(h["lr"]) * d).astype(F)); m_out.append(buf.astype(F))
    upd = ema_updates + 1
    dec = F(ema_decay * (1 - math.exp(-upd / ema_tau)))
    e_out = [(dec * e + (F(1) - dec) * p).astype(F) for e, p in zip(emas, p_out)]
    eb_out = [(dec * e + (F(1) - dec) * b).astype(F) for e, b in zip(ema_buffers, buffers)]
    return p_out, m_out, e_out, eb_out, float(total), not finite


def synth_problem(seed: int, shapes=((16, 8, 3, 3), (16,), (16,), (32, 16, 1, 1), (17000,), (17,)), groups=(1, 2, 0, 1, 1, 0)):
    """Seeded parameter / gradient / momentum / EMA tensors (float32) for the step above."""
    rs = np.random.RandomState(seed)
    params = [rs.normal(0, 0.5, s).astype(F) for s in shapes]
    grads = [rs.normal(0, 2.0, s).astype(F) for s in shapes]
    moms = [rs.normal(0, 0.3, s).astype(F) for s in shapes]
    emas = [(p + rs.normal(0, 0.01, p.shape)).astype(F) for p in params]
    return params, grads, moms, emas, list(groups)


FIXTURE_STRIDE = 5  # tests/golden/optim.npz keeps every 5th element of each result (flattened) to stay small
