"""Training step of the density estimators on the hot path (loss -> backward -> gradient all-reduce -> optimiser).

The reference keeps no training module (``torch_nf/lfi.py`` is missing from the repository); its loops live in the
notebooks and scripts: ``notebooks/LFI_learning_rules.ipynb:295-306`` (``loss = -mean(cnf.log_prob(z[:, None, :], x))``,
``Adam(lr=1e-4)``, ``zero_grad / backward / step``) with the shapes of ``scripts/lfi_mat.py:35-46``.  The same step
here: the forward and backward of every bijector are the CUDA kernels behind ``torch.autograd.Function``s
(activations recomputed from the saved layer input), the hyper-network is plain torch, and - one process per GPU, the
context rows sharded - the gradients are summed over the ranks in ONE flat bucket (NCCL) before the replicated
optimiser step, and the loss is averaged.
"""
import torch

from . import dist


def nde_loss(cde, z, x):
    """-mean log q(z | x): ``z (M, D)`` or ``(M, N, D)`` observations, ``x (M, D_x)`` contexts
    (notebooks/LFI_learning_rules.ipynb:299-301)."""
    if z.dim() == 2:
        z = z[:, None, :]
    return -torch.mean(cde.log_prob(z, x))


def mle_loss(nf, z):
    """-mean log q(z) of an unconditional flow with its own parameters (``nf.params``)."""
    return -torch.mean(nf.log_prob(z))


def allreduce_gradients(parameters):
    """Average the gradients over the ranks: the local losses are means over equally sized shards, so the global
    gradient is the mean of the local ones.  One flat bucket, one collective."""
    grads = [p.grad for p in parameters if p.grad is not None]
    if dist.is_enabled() and dist.world_size() > 1 and grads:
        dist.allreduce_sum_(grads)
        w = float(dist.world_size())
        for g in grads:
            g.div_(w)
    return grads


def train_step(loss_fn, parameters, optimizer):
    """One optimiser step; returns the (rank-averaged) loss as a 0-dim tensor on the device, without a host sync."""
    parameters = list(parameters)
    optimizer.zero_grad(set_to_none=True)
    loss = loss_fn()
    loss.backward()
    allreduce_gradients(parameters)
    optimizer.step()
    return dist.allreduce_mean_scalar(loss.detach())
