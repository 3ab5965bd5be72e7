"""Training loops over the hot path: SNPE / APT posterior estimation and EFN training.

The reference's ``torch_nf/lfi.py`` is MISSING from its repository (``scripts/lfi_mat.py:5`` imports ``train_APT`` from
it); what survives are the notebook loops it was factored out of.  They are restated here over ``train.train_step``
(loss -> backward -> one gradient all-reduce -> optimiser):
  * ``train_SNPE``: ``notebooks/LFI_learning_rules.ipynb:274-306`` (``train_nde``): rounds of proposals (round 1 the
    prior, later rounds the current posterior estimate ``cnf(x0, N)``), simulate, ``loss = -mean(cnf.log_prob(z[:, None, :], x))``,
    Adam(lr 1e-4), gradient clamp, step.
  * ``train_APT``: the call of ``scripts/lfi_mat.py:48-57`` (``M`` contexts per step, ``M_atom`` atoms, ``R`` rounds,
    returns ``cnf, losses, zs, log_probs, it_time``) with the atomic-proposal loss of Greenberg et al. 2019 (APT), the
    algorithm that script names: per context m the atoms are z_m and ``M_atom - 1`` other parameters of the batch,
    ``loss = -mean_m log( q(z_m | x_m) / p(z_m) / sum_a q(z_a | x_m) / p(z_a) )``.  One ``cnf.log_prob`` call on
    ``z (M, M_atom, D)`` per step - the "several samples per context" regime of the kernels.  Parity with the lost file
    cannot be pinned; the loss is tested against its closed form instead (tests/test_lfi_host.py).
  * ``EFNLoss`` / ``train_efn``: ``notebooks/two_network_arch.ipynb:77-112``: ``mean(log_q - eta . T(z))`` through the
    SAMPLE direction of the flow.
A ``system`` is the notebooks' ``System`` protocol (``LFI_learning_rules.ipynb:203-213``): ``D``, ``sample_prior(M) ->
(z (M, D), p_z (M,))``, ``simulate(z) -> x (M, D_x)`` (numpy in, numpy out), optional ``support_layer``.
"""
import time

import numpy as np
import torch

from . import train


def clip_grads(parameters, clip):
    """Clamp every gradient element to [-clip, clip] (LFI_learning_rules.ipynb:268-271)."""
    for p in parameters:
        if p.grad is not None:
            p.grad.data.clamp_(-clip, clip)


def _as_float_tensor(a, device=None):
    t = a if isinstance(a, torch.Tensor) else torch.tensor(np.asarray(a))
    t = t.float()
    return t if device is None else t.to(device)


def _device_of(cnf):
    for p in cnf.parameters():
        return p.device
    return torch.device("cpu")


def _propose(r, M, cnf, system, x0):
    """Round-r proposals: the prior in round 1, the current posterior estimate at the observation afterwards
    (``SNPE_prior``, LFI_learning_rules.ipynb:248-266).  Returns z (M, D) and the proposal density at z."""
    if r == 1:
        z, p_z = system.sample_prior(M)
        return _as_float_tensor(z), _as_float_tensor(p_z)
    with torch.no_grad():
        z, log_q_z = cnf(x0, N=M)
    return z[0].detach().float().cpu(), torch.exp(log_q_z[0].detach().float()).cpu()


def snpe_loss(cnf, z, x):
    """-mean log q(z | x) (LFI_learning_rules.ipynb:299-301)."""
    return train.nde_loss(cnf, z, x)


def apt_atoms(M, M_atom, generator=None):
    """Atom indices (M, M_atom): column 0 is the context's own parameter, the others are M_atom - 1 DISTINCT other rows
    of the batch, uniformly at random."""
    if M_atom < 2 or M_atom > M:
        raise ValueError("APT needs 2 <= M_atom <= M, got M_atom=%d, M=%d" % (M_atom, M))
    scores = torch.rand(M, M, generator=generator)
    scores.fill_diagonal_(2.0)                         # never picked among the "others"
    others = scores.argsort(dim=1)[:, :M_atom - 1]
    return torch.cat((torch.arange(M)[:, None], others), dim=1)


def apt_loss(cnf, z, x, M_atom, log_prior=None, atoms=None, generator=None):
    """Atomic-proposal posterior loss.  ``z (M, D)``, ``x (M, D_x)``; ``log_prior(z (..., D)) -> (...)`` (None: a flat
    prior over the flow's support, the ratio cancels).  One log_prob call on (M, M_atom, D)."""
    M = z.shape[0]
    idx = apt_atoms(M, M_atom, generator) if atoms is None else atoms
    za = z[idx.to(z.device)]                           # (M, M_atom, D)
    lp = cnf.log_prob(za, x)                           # (M, M_atom)
    if log_prior is not None:
        lp = lp - log_prior(za).to(lp.dtype)
    return -torch.mean(lp[:, 0] - torch.logsumexp(lp, dim=1))


def _run_rounds(cnf, system, x0, loss_of, M, R, num_iters, lr, clip, verbose):
    dev = _device_of(cnf)
    x0_t = _as_float_tensor(x0, dev)
    params = list(cnf.parameters())
    opt = torch.optim.Adam(params, lr=lr)
    losses, zs, log_probs = [], [], []
    t0 = time.time()
    n_it = 0
    for r in range(1, R + 1):
        for i in range(1, num_iters + 1):
            z, _ = _propose(r, M, cnf, system, x0_t)
            x = _as_float_tensor(system.simulate(z.numpy()), dev)
            zd = z.to(dev)

            def step_loss():
                return loss_of(zd, x)

            def clipped_step():
                # train_step with the notebook's gradient clamp between backward and the optimiser
                opt.zero_grad(set_to_none=True)
                loss = step_loss()
                loss.backward()
                clip_grads(params, clip)
                train.allreduce_gradients(params)
                opt.step()
                return loss.detach()
            loss = float(clipped_step())
            n_it += 1
            losses.append(loss)
            if verbose and (i == 1 or i % 10 == 0):
                print("round %d it %d, loss=%.2E" % (r, i, loss))
            if not np.isfinite(loss):
                break
        with torch.no_grad():                           # the round's posterior estimate at the observation
            z_r, lq_r = cnf(x0_t, N=M)
        zs.append(z_r[0].detach().float().cpu().numpy())
        log_probs.append(lq_r[0].detach().float().cpu().numpy())
    it_time = (time.time() - t0) / max(n_it, 1)
    return cnf, np.array(losses), np.array(zs), np.array(log_probs), it_time


def train_SNPE(cnf, system, x0, M=500, R=4, num_iters=1000, lr=1e-4, clip=1e10, verbose=False):
    """Sequential neural posterior estimation with the plain conditional-density loss (``train_nde``)."""
    return _run_rounds(cnf, system, x0, lambda z, x: snpe_loss(cnf, z, x), M, R, num_iters, lr, clip, verbose)


def train_APT(cnf, system, x0, M=2000, M_atom=100, R=6, num_iters=5000, lr=1e-4, clip=1e10, log_prior=None,
              verbose=False):
    """Automatic posterior transformation with atomic proposals (signature of scripts/lfi_mat.py:48-57)."""
    return _run_rounds(cnf, system, x0, lambda z, x: apt_loss(cnf, z, x, M_atom, log_prior), M, R, num_iters, lr, clip,
                       verbose)


def EFNLoss(z, log_prob, eta, T):
    """mean(log q(z) - eta . T(z)) over contexts and samples (two_network_arch.ipynb:77-83).  ``z (M, N, D)``,
    ``log_prob (M, N)``, ``eta (M, D_eta)``."""
    eta_dot_T = torch.matmul(T(z), eta[:, :, None].to(z.dtype))[:, :, 0]
    return torch.mean(log_prob.to(eta_dot_T.dtype) - eta_dot_T)


def train_efn(cnf, exp_fam, num_iters=1000, M=50, N=100, lr=1e-4, verbose=False):
    """Exponential-family network training (two_network_arch.ipynb:85-112): sample eta, sample the flow conditioned on
    it, minimise the EFN loss.  Returns (losses, KLs)."""
    dev = _device_of(cnf)
    params = list(cnf.parameters())
    opt = torch.optim.Adam(params, lr=lr)
    losses, KLs = [], []
    for i in range(1, num_iters + 1):
        eta_np = exp_fam.sample_eta(N=M)
        eta = _as_float_tensor(eta_np, dev)
        out = {}

        def loss_fn():
            z, log_prob = cnf(eta, N=N)
            out["z"], out["lp"] = z, log_prob
            return EFNLoss(z, log_prob, eta, exp_fam.T)
        loss = float(train.train_step(loss_fn, params, opt))
        if not np.isfinite(loss):
            break
        KL = float(np.mean(exp_fam.KL(out["z"].detach().float().cpu().numpy(), out["lp"].detach().float().cpu().numpy(), eta_np)))
        if verbose and (i == 1 or i % 100 == 0):
            print("%d: loss=%.2E, KL=%.2E" % (i, loss, KL))
        losses.append(loss)
        KLs.append(KL)
    return losses, KLs
