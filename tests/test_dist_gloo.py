"""Host-side data-parallel logic on CPU: world_size 2, gloo backend.

The hot path shards by sample rows; the only exchange steps are the BatchNorm
batch statistics (float64 [sum | sumsq | rows] buffers, as produced by
tnf_colstats) and the gradient / scalar reductions.  Here each rank plays a
shard with the CPU oracle and the collectives go through torch_nf_b200.dist."""
import os
import socket

import numpy as np
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from oracle import flow_oracle as O
from torch_nf_b200 import dist, train
from torch_nf_b200.synthetic import synthetic_params


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _colstats_like_kernel(z):
    """[sum (D) | sumsq (D) | rows] in float64: the layout tnf_colstats writes."""
    zv = z.reshape(-1, z.shape[-1]).double()
    return torch.cat([zv.sum(0), (zv * zv).sum(0), torch.tensor([float(zv.shape[0])], dtype=torch.float64)])


def _finalize_like_kernel(sums, D, eps):
    n = sums[2 * D]
    mean = sums[:D] / n
    var = (sums[D:2 * D] / n - mean * mean).clamp_min(0)
    return mean.float(), torch.sqrt(var + eps).float()


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dist.enable()
        assert dist.world_size() == world and dist.rank() == rank
        res = {}
        # 1. contiguous shards cover the batch exactly once, sizes differ by at most one
        total = 1001
        lo, hi = dist.shard_range(total)
        spans = [dist.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
        # 2. BatchNorm statistics of the GLOBAL batch from sharded partial sums
        D = 8
        rs = np.random.RandomState(0)
        z_full = torch.tensor((rs.standard_normal((1, total, D)) * 3.0 + 5.0).astype(np.float32))
        z_shard = z_full[:, lo:hi]
        sums = dist.allreduce_stats(_colstats_like_kernel(z_shard))
        mean, alpha = _finalize_like_kernel(sums, D, 1e-5)
        _, _, mean_ref, alpha_ref = O.batchnorm_forward(z_full, 1e-5)
        res["bn_mean_err"] = float((mean - mean_ref).abs().max())
        res["bn_alpha_err"] = float((alpha - alpha_ref).abs().max())
        res["rows"] = float(sums[2 * D])
        # 3. sharded sample + log_prob of a whole flow equals the single-process result
        chain = O.build_chain(D, "coupling", 2, 2, 15)
        spec = [(b["kind"], b.get("L", 0), b.get("U", 0), b.get("upper", False)) for b in chain]
        params = torch.tensor(synthetic_params(spec, D, 1, seed=3))
        omega = rs.standard_normal((1, total, D))
        z_ref, lq_ref, st_ref = O.normflow_forward(chain, D, params, omega)
        # sharded forward: every BatchNorm all-reduces its statistics
        z = torch.tensor(omega[:, lo:hi]).float()
        lq = torch.tensor(O.base_log_density_f64(omega[:, lo:hi]))
        idx, st = 0, []
        for b in chain:
            if b["kind"] == "BatchNorm":
                s = dist.allreduce_stats(_colstats_like_kernel(z))
                m, a = _finalize_like_kernel(s, D, b["eps"])
                z, ld, _, _ = O.batchnorm_forward(z, b["eps"], True, m, a)
                st.append((m, a))
            elif b["kind"] == "RealNVP":
                n = O.bijector_num_params(b, D)
                z, ld = O.coupling_forward(z, params[:, idx:idx + n], D, b["L"], b["U"], b["upper"]); idx += n
                st.append(None)
            else:
                z, ld = O.affine_forward(z, params[:, idx:idx + 2 * D], D); idx += 2 * D
                st.append(None)
            lq = lq - ld
        res["z_err"] = float((z - z_ref[:, lo:hi]).abs().max())
        res["lq_err"] = float((lq - lq_ref[:, lo:hi]).abs().max())
        # log_prob needs no communication once the statistics are shared
        lp = O.normflow_log_prob(chain, D, z_ref[:, lo:hi], params, st)
        lp_ref = O.normflow_log_prob(chain, D, z_ref, params, st_ref)
        res["lp_err"] = float((lp - lp_ref[:, lo:hi]).abs().max())
        # 4. gradient all-reduce (sum) and scalar mean
        p = params.clone().requires_grad_(True)
        loss = -O.normflow_log_prob(chain, D, z_ref[:, lo:hi], p, st_ref).sum() / total
        loss.backward()
        g = p.grad.clone()
        dist.allreduce_sum_([g])
        p2 = params.clone().requires_grad_(True)
        (-O.normflow_log_prob(chain, D, z_ref, p2, st_ref).mean()).backward()
        res["grad_rel"] = float((g - p2.grad).norm() / p2.grad.norm())
        m = dist.allreduce_mean_scalar(torch.tensor(float(rank + 1)))
        res["mean_scalar"] = float(m)
        # 5. train.train_step on equally sized shards: the averaged local gradients are the full-batch gradient, the
        #    replicated optimiser step leaves identical parameters on every rank
        even = (total // world) * world
        lo_e, hi_e = dist.shard_range(even)
        pw = torch.nn.Parameter(params.clone())
        opt = torch.optim.Adam([pw], lr=1e-3)
        loss = train.train_step(lambda: -O.normflow_log_prob(chain, D, z_ref[:, lo_e:hi_e], pw, st_ref).mean(), [pw], opt)
        pf = torch.nn.Parameter(params.clone())
        optf = torch.optim.Adam([pf], lr=1e-3)
        optf.zero_grad()
        loss_f = -O.normflow_log_prob(chain, D, z_ref[:, :even], pf, st_ref).mean()
        loss_f.backward()
        optf.step()
        res["train_loss_err"] = abs(float(loss) - float(loss_f))
        res["train_param_err"] = float((pw.detach() - pf.detach()).abs().max())
        if rank == 0:
            out.put(res)
    finally:
        dist.disable()
        td.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["rows"] == 1001.0
    assert res["bn_mean_err"] < 1e-5 and res["bn_alpha_err"] < 1e-5
    assert res["z_err"] < 2e-5 and res["lq_err"] < 1e-4 and res["lp_err"] < 1e-4
    assert res["grad_rel"] < 1e-5
    assert res["mean_scalar"] == 1.5
    assert res["train_loss_err"] < 1e-4 and res["train_param_err"] < 1e-6


def test_single_process_defaults():
    assert not dist.is_enabled() and dist.world_size() == 1 and dist.rank() == 0
    assert dist.shard_range(10) == (0, 10)
    s = torch.arange(5, dtype=torch.float64)
    assert dist.allreduce_stats(s) is s


def test_peer_struct_sequence_numbers():
    """tnf_peer_t as the host builds it: every call gets the current sequence number and advances the counter by the
    number of exchanges it may perform, so all ranks (which make the same calls) stay in step."""
    from torch_nf_b200 import _lib
    assert dist.peer_struct(3) is None                      # peer exchange not set up
    dist._peer = dict(stats_ptrs=[0x1000, 0x2000], flag_ptrs=[0x3000, 0x4000], seq=1)
    try:
        p = dist.peer_struct(8)
        assert isinstance(p, _lib.Peer) and p.seq == 1 and p.world == 1 and p.rank == 0     # no process group here
        assert p.stats[0] == 0x1000 and p.flags[0] == 0x3000
        assert dist.peer_struct(8).seq == 9 and dist.peer_struct(0).seq == 17 and dist.peer_struct(2).seq == 18
    finally:
        dist._peer = None
