"""world_size-2 gloo test (CPU) of the data-parallel train-step decomposition used by
``ops.MlpTrainer``: each rank holds a shard of the rows; one all-reduce of the five
statistic sums and one of the flat gradient reproduce the single-process global-batch step."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import wvn_path


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, shards, sd, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, y, yv = shards[rank]
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    res = wvn_path.mlp_forward(x, params)
    D = x.shape[1]
    loss_reco = ((res[:, 1:] - x) ** 2).mean(1)
    raw = (res[:, 0] - y) ** 2
    # phase 1: local sums -> all-reduce (what wvn_mlp_train_forward_stats leaves in scalars[0..4])
    lr_v = loss_reco.detach()[yv].double()
    sums = torch.stack([lr_v.sum(), (lr_v**2).sum(), raw.detach().double().sum(), yv.double().sum(),
                        torch.tensor(float(x.shape[0]), dtype=torch.float64)])
    dist.all_reduce(sums)
    n_valid, n_total = sums[3].item(), int(sums[4].item())
    mean = (sums[0] / n_valid).float().reshape(1)
    std = (((sums[1] - n_valid * (sums[0] / n_valid) ** 2) / (n_valid - 1)).clamp_min(0).sqrt()).float().reshape(1)
    conf = wvn_path.confidence_inference(loss_reco.detach(), mean, std, 0.5)
    # phase 2: local loss contribution with GLOBAL normalisers -> backward -> all-reduce of the grads
    w = torch.where(yv, torch.ones_like(conf), 1 - conf)
    local = 0.03 * (raw * w).sum() / n_total + 0.5 * loss_reco[yv].sum() / n_valid
    local.backward()
    flat = torch.cat([p.grad.reshape(-1) for p in params.values()])
    dist.all_reduce(flat)
    if rank == 0:
        ret["grads"], ret["mean"], ret["std"] = flat, mean, std
    dist.destroy_process_group()


def test_two_rank_step_equals_global_batch_step():
    torch.manual_seed(0)
    D, R = 32, 60
    sd = wvn_path.mlp_init(D, (16, 8), seed=42)
    x = torch.randn(R, D)
    y, yv = wvn_path.synthetic_supervision(R, seed=3, p_valid=0.3)
    _, _, ref = wvn_path.train_step(sd, None, x, y, yv)
    shards = [(x[:25], y[:25], yv[:25]), (x[25:], y[25:], yv[25:])]  # ragged shards
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), shards, sd, ret), nprocs=2, join=True)
    want = torch.cat([ref["grads"][k].reshape(-1) for k in sd])
    assert (ret["grads"] - want).abs().max() < 1e-6
    assert abs(ret["mean"].item() - ref["mean"]) < 1e-6 and abs(ret["std"].item() - ref["std"]) < 1e-6
