"""Generate tests/golden/mod_*.npz from the REFERENCE's own modules (run in the build container only).

    NUMBA_DISABLE_CUDA=1 NUMBA_CACHE_DIR=/tmp/nbcache PYTHONPATH=/root/reference python oracle/gen_golden_modules.py

Pins the device-resident restatements in flash-attention-minitorch_b200/fa_b200/mt/ 
against the reference's Linear, LayerNorm1d, GELU, softmax_loss, Embedding, FeedForward and Adam, executed
unmodified on minitorch's numba CPU backend (2-D / 3-D shapes only: its 4-D matmul is wrong on that backend,
SURVEY.md §8c).  Nothing is copied; outputs and gradients are recorded for seeded inputs.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden import OUT, fastops_backend  # noqa: E402


def main():
    minitorch, be = fastops_backend()
    from minitorch.modules_basic import Embedding, LayerNorm1d, Linear
    from minitorch.modules_transfomer import FeedForward
    rng = np.random.default_rng(5001)
    t = lambda a, g=False: minitorch.tensor_from_numpy(np.ascontiguousarray(a, dtype=np.float32), backend=be, requires_grad=g)  # noqa: E731
    out = {}

    # Linear (modules_basic.py:93-122)
    n, i, o = 12, 16, 24
    x = rng.uniform(-1, 1, (n, i)).astype(np.float32)
    gy = rng.uniform(-1, 1, (n, o)).astype(np.float32)
    lin = Linear(i, o, True, be)
    W, b = lin.weights.value.to_numpy().copy(), lin.bias.value.to_numpy().copy()
    tx = t(x, True)
    y = lin(tx)
    (y * t(gy)).sum().backward()
    out.update(lin_x=x, lin_w=W, lin_b=b, lin_gy=gy, lin_y=y.to_numpy(), lin_dx=tx.grad.to_numpy(),
               lin_dw=lin.weights.value.grad.to_numpy(), lin_db=lin.bias.value.grad.to_numpy())

    # LayerNorm1d (modules_basic.py:143-175), non-trivial weight / bias
    n, dim, eps = 10, 32, 1e-5
    x = rng.normal(0, 2, (n, dim)).astype(np.float32)
    gy = rng.uniform(-1, 1, (n, dim)).astype(np.float32)
    ln = LayerNorm1d(dim, eps, be)
    w, b = rng.uniform(0.5, 1.5, dim).astype(np.float32), rng.uniform(-0.5, 0.5, dim).astype(np.float32)
    ln.weights.update(t(w, True))
    ln.bias.update(t(b, True))
    tx = t(x, True)
    y = ln(tx)
    (y * t(gy)).sum().backward()
    out.update(ln_x=x, ln_w=w, ln_b=b, ln_eps=np.float32(eps), ln_gy=gy, ln_y=y.to_numpy(), ln_dx=tx.grad.to_numpy(),
               ln_dw=ln.weights.value.grad.to_numpy(), ln_db=ln.bias.value.grad.to_numpy())

    # GELU (nn.py:208-214)
    x = rng.normal(0, 2, (6, 20)).astype(np.float32)
    gy = rng.uniform(-1, 1, (6, 20)).astype(np.float32)
    tx = t(x, True)
    y = minitorch.nn.GELU(tx)
    (y * t(gy)).sum().backward()
    out.update(gelu_x=x, gelu_gy=gy, gelu_y=y.to_numpy(), gelu_dx=tx.grad.to_numpy())

    # softmax_loss (nn.py:256-274)
    n, C = 9, 37
    logits = rng.normal(0, 3, (n, C)).astype(np.float32)
    target = rng.integers(0, C, n)
    gw = rng.uniform(0, 1, n).astype(np.float32)
    tl = t(logits, True)
    loss = minitorch.nn.softmax_loss(tl, minitorch.tensor_from_numpy(target.astype(np.float32), backend=be))
    (loss * t(gw)).sum().backward()
    out.update(sl_logits=logits, sl_target=target.astype(np.int64), sl_gw=gw, sl_loss=loss.to_numpy(), sl_dlogits=tl.grad.to_numpy())

    # Embedding (modules_basic.py:17-53): one_hot @ weights
    V, E, bs, T = 23, 8, 3, 5
    emb = Embedding(V, E, be)
    We = emb.weights.value.to_numpy().copy()
    idx = rng.integers(0, V, (bs, T))
    ge = rng.uniform(-1, 1, (bs, T, E)).astype(np.float32)
    y = emb(minitorch.tensor_from_numpy(idx.astype(np.float32), backend=be))
    (y * t(ge)).sum().backward()
    out.update(emb_w=We, emb_idx=idx.astype(np.int64), emb_gy=ge, emb_y=y.to_numpy(), emb_dw=emb.weights.value.grad.to_numpy())

    # FeedForward (modules_transfomer.py:160-201) with dropout disabled
    bs, T, E, mid = 2, 6, 16, 24
    ff = FeedForward(E, mid, 0.0, True, be)
    ff.eval()
    x = rng.uniform(-1, 1, (bs, T, E)).astype(np.float32)
    gy = rng.uniform(-1, 1, (bs, T, E)).astype(np.float32)
    tx = t(x, True)
    y = ff(tx)
    (y * t(gy)).sum().backward()
    out.update(ff_x=x, ff_gy=gy, ff_y=y.to_numpy(), ff_dx=tx.grad.to_numpy(),
               ff_w_in=ff.linear_in.weights.value.to_numpy(), ff_b_in=ff.linear_in.bias.value.to_numpy(),
               ff_w_out=ff.linear_out.weights.value.to_numpy(), ff_b_out=ff.linear_out.bias.value.to_numpy(),
               ff_dw_in=ff.linear_in.weights.value.grad.to_numpy(), ff_dw_out=ff.linear_out.weights.value.grad.to_numpy())

    # Adam (optim.py:30-76): three steps with prescribed gradients
    p0 = rng.uniform(-1, 1, (4, 5)).astype(np.float32)
    grads = rng.normal(0, 1, (3, 4, 5)).astype(np.float32)
    par = minitorch.Parameter(t(p0, True))
    opt = minitorch.Adam([par], lr=0.02)
    traj = []
    for g in grads:
        par.value.grad = t(g)  # what backward() would have left there
        opt.step()
        traj.append(par.value.to_numpy().copy())
    out.update(adam_p0=p0, adam_grads=grads, adam_traj=np.stack(traj), adam_lr=np.float32(0.02))

    np.savez_compressed(os.path.join(OUT, "mod_reference_modules.npz"),
                        source=np.array("reference minitorch modules on FastOps (numba CPU), fp32"),
                        **{k: np.asarray(v) for k, v in out.items()})
    print("wrote mod_reference_modules.npz with", len(out), "arrays")


if __name__ == "__main__":
    sys.exit(main())
