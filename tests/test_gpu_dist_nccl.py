"""Multi-rank parity on GPUs: 2 NCCL ranks, `omega` sharded by rows, BatchNorm statistics all-reduced
(torch_nf/bijectors.py:401-415 pools over the WHOLE batch) must reproduce the single-process golden vectors of the
unmodified reference.  Needs 2 visible GPUs (skipped otherwise): run with `gpurun --gpus 2`."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, precision, peer, q):
    import torch.distributed as td
    import torch_nf_b200 as tnf
    import torch_nf_b200.density_estimator as de
    from torch_nf_b200 import config, dist
    from torch_nf_b200.synthetic import chain_spec, synthetic_params
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    td.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        dist.enable()
        peer_on = dist.enable_peer_exchange() if peer else False
        tnf.set_conditioner_precision(precision)
        config.set_tc_min_rows(1)
        g = np.load(os.path.join(GOLDEN, "flow_c3.npz"))
        D, stages, L, U, M, N, pseed, oseed = [int(v) for v in g["cfg"]]
        nf = de.NormFlow(D, True, "coupling", stages, L, U)
        params = torch.tensor(synthetic_params(chain_spec(nf.bijectors), D, M, seed=pseed)).cuda()
        np.random.seed(oseed)
        omega = np.random.normal(0.0, 1.0, (M, N, D))
        lo, hi = dist.shard_range(N)
        with torch.no_grad():
            z, lq = nf.forward(params, hi - lo, omega=omega[:, lo:hi])
            lp = nf.log_prob(torch.tensor(g["z"][:, lo:hi]).cuda(), params)
        bn = [b for b in nf.bijectors if b.name == "BatchNorm"]
        if peer_on:      # a second call: the sequence counters and the slot parity carry over
            with torch.no_grad():
                z2, lq2 = nf.forward(params, hi - lo, omega=omega[:, lo:hi])
            assert torch.equal(z, z2) and torch.equal(lq, lq2)
        q.put((rank, lo, hi, z.cpu().numpy(), lq.cpu().numpy(), lp.cpu().numpy(),
               [b.get_last_mean().cpu().numpy() for b in bn], [b.get_last_alpha().cpu().numpy() for b in bn],
               peer_on, dist.peer_error))
    finally:
        td.destroy_process_group()


@pytest.mark.parametrize("precision,peer", [("fp32", False), ("bf16", False), ("fp32", True), ("bf16", True)])
def test_two_nccl_ranks_reproduce_the_single_process_golden(precision, peer):
    """peer = True: the statistics cross the ranks inside the fold kernel over NVLink peer memory (tnf_peer_t) instead
    of one NCCL all-reduce per BatchNorm; both ranks must then hold bit-identical statistics."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from torch_nf_b200 import config
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, precision, peer, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if peer:
        if not all(r[8] for r in res):
            pytest.skip("symmetric memory unavailable: %s" % res[0][9])
        for i in range(len(res[0][6])):     # summed in rank order on every rank: identical bits
            assert np.array_equal(res[0][6][i], res[1][6][i]) and np.array_equal(res[0][7][i], res[1][7][i])
    g = np.load(os.path.join(GOLDEN, "flow_c3.npz"))
    z = np.concatenate([r[3] for r in res], axis=1)
    lq = np.concatenate([r[4] for r in res], axis=1)
    lp = np.concatenate([r[5] for r in res], axis=1)
    rel = lambda a, b: float((np.abs(a - b) / np.maximum(1.0, np.abs(b))).max())
    dz = float(np.abs(z - g["z"]).max())
    print("2 ranks, %s: max|dz| %.3g rel z %.3g rel log_q %.3g rel log_prob %.3g" % (precision, dz, rel(z, g["z"]), rel(lq, g["log_q_z"]), rel(lp, g["log_prob"])))
    if precision == "fp32":
        assert rel(z, g["z"]) <= config.FP32_TOL_Z and rel(lq, g["log_q_z"]) <= config.FP32_TOL_LOGP
        assert rel(lp, g["log_prob"]) <= config.FP32_TOL_LOGP
    else:
        assert dz <= config.BF16_TOL_Z and rel(lq, g["log_q_z"]) <= config.BF16_TOL_LOGP
        assert rel(lp, g["log_prob"]) <= config.BF16_TOL_LOGP
    for r in res:      # every rank holds the GLOBAL batch statistics
        for i in range(len(r[6])):
            np.testing.assert_allclose(r[6][i], g["bn_mean"][i], rtol=1e-4 if precision == "fp32" else 5e-2, atol=2e-5 if precision == "fp32" else 5e-2)
            np.testing.assert_allclose(r[7][i], g["bn_alpha"][i], rtol=1e-4 if precision == "fp32" else 5e-2, atol=1e-6 if precision == "fp32" else 5e-2)
