"""a15: the eikonal term as the build defines it — oracle only this round (the CUDA path rejects eikonal_weight > 0).
Checks of the restatement itself: normals against central differences inside a cell, the loss on a table that encodes an exact
distance-like field, and the double-backward parameter gradient against finite differences in float64."""
import numpy as np
import torch

from oracle import nof_oracle as O


def _params(L=3, finest=32, log2T=10, seed=0, dtype=torch.float64, scale=0.3):
    offsets, pls = O.grid_offsets(L, 16, finest, log2T)
    g = torch.Generator().manual_seed(seed)
    P = {'embeddings': ((torch.rand(int(offsets[-1]), 2, generator=g, dtype=dtype) * 2 - 1) * scale).requires_grad_(True),
         'offsets': offsets, 'S': float(np.log2(pls)), 'H': 16}
    for k, v in O.init_mlp(L * 2, 9, seed=seed).items():
        P[k] = v.to(dtype).requires_grad_(True)
    return P


def test_normals_match_central_differences_inside_cells():
    P = _params()
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(64, 3, generator=g, dtype=torch.float64) * 1.6 - 0.8)
    valid = torch.ones(64, dtype=torch.bool)
    sdf, n = O.sdf_normals(P, x, valid)
    h = 1e-6
    fd = torch.zeros_like(n)
    for d in range(3):
        e = torch.zeros(3, dtype=torch.float64); e[d] = h
        sp, _ = O.sdf_normals(P, x + e, valid)
        sm, _ = O.sdf_normals(P, x - e, valid)
        fd[:, d] = ((sp - sm) / (2 * h)).detach()
    # trilinear interpolation + ReLU are piecewise smooth: compare where the +-h stencil stays in one piece
    ok = (fd - n.detach()).abs().max(dim=1).values < 1e-4
    assert ok.float().mean() > 0.9
    np.testing.assert_allclose(n.detach()[ok].numpy(), fd[ok].numpy(), atol=1e-4)


def test_invalid_samples_count_as_unit_error_and_zero_normal():
    P = _params()
    x = torch.tensor([[0.1, 0.2, 0.3], [1.5, 0.0, 0.0]], dtype=torch.float64)         # the second one is outside the box
    valid = (x.abs() <= 1).all(dim=-1)
    sdf, n = O.sdf_normals(P, x, valid)
    assert sdf[1] == 0 and bool((n[1] == 0).all())
    one = O.eikonal_loss(P, x[1:], valid[1:], 0.5)
    assert abs(float(one.detach()) - 0.5) < 1e-12                                              # (|0| - 1)^2 = 1


def test_eikonal_gradient_matches_finite_differences():
    P = _params(L=2, finest=24, log2T=8, seed=3)
    g = torch.Generator().manual_seed(2)
    x = (torch.rand(24, 3, generator=g, dtype=torch.float64) * 1.6 - 0.8)
    valid = torch.ones(24, dtype=torch.bool)
    loss = O.eikonal_loss(P, x, valid, 0.1)
    leaves = {'embeddings': P['embeddings'], 'w1': P['sigma_net.0.weight'], 'w2': P['sigma_net.2.weight']}
    grads = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
    assert float(loss) > 0
    rng = np.random.default_rng(0)
    for (name, p), gr in zip(leaves.items(), grads):
        assert gr is not None and float(gr.abs().max()) > 0, name
        flat = p.detach().view(-1)
        nz = gr.view(-1).abs().argsort(descending=True)[:200].numpy()
        checked = 0
        for i in rng.choice(nz, size=6, replace=False):
            h = 1e-6
            old = float(flat[i])
            with torch.no_grad():
                flat[i] = old + h
            lp = float(O.eikonal_loss(P, x, valid, 0.1))
            with torch.no_grad():
                flat[i] = old - h
            lm = float(O.eikonal_loss(P, x, valid, 0.1))
            with torch.no_grad():
                flat[i] = old
            fd = (lp - lm) / (2 * h)
            an = float(gr.view(-1)[i])
            if abs(fd - an) <= 2e-3 * max(abs(an), 1e-6):
                checked += 1
        assert checked >= 5, (name, checked)            # a perturbation may flip a ReLU / the sdf<1 selection: allow one miss
    # sigma_net biases only enter through the ReLU masks: no gradient path
    gb = torch.autograd.grad(O.eikonal_loss(P, x, valid, 0.1), [P['sigma_net.0.bias']], allow_unused=True)[0]
    assert gb is None or float(gb.abs().max()) == 0.0
