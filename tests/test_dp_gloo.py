"""CPU, world_size 2, gloo: the data-parallel plumbing of cocosnet_b200.trainer
(contiguous batch shard + bucketed gradient all-reduce) gives the same averaged
gradients and the same parameters after k steps as a single process on the
whole batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _model():
    torch.manual_seed(3)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.LeakyReLU(0.2),
                               torch.nn.Conv2d(8, 4, 3, padding=1), torch.nn.Flatten(), torch.nn.Linear(4 * 64, 5))


def _batch():
    g = torch.Generator().manual_seed(9)
    return {"image": torch.randn(8, 3, 8, 8, generator=g), "target": torch.randn(8, 5, generator=g),
            "path": ["p%d" % i for i in range(8)]}


def _train(model, data, steps, world):
    from cocosnet_b200 import trainer as tr
    opt = torch.optim.Adam(model.parameters(), lr=1e-2, betas=(0.0, 0.9))
    tr._BUCKET_BYTES = 1024  # force several buckets
    for _ in range(steps):
        opt.zero_grad()
        d = tr.shard_batch(data)
        loss = (model(d["image"]) - d["target"]).pow(2).mean()  # per-replica mean, like the reference losses
        loss.backward()
        tr.allreduce_grads(list(model.parameters()))
        opt.step()
    return [p.detach().clone() for p in model.parameters()]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        params = _train(_model(), _batch(), 3, world)
        torch.save(params, os.path.join(out_dir, "rank%d.pt" % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_match_single_process(tmp_path):
    ref = _train(_model(), _batch(), 3, 1)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"))
    for a, b in zip(r0, r1):
        assert torch.equal(a, b), "replicas diverged"
    for a, b in zip(r0, ref):
        assert torch.allclose(a, b, atol=1e-6), float((a - b).abs().max())


def test_shard_batch_is_contiguous():
    from cocosnet_b200 import trainer as tr
    d = _batch()
    s1 = tr.shard_batch(d, rank=1, world=2)
    assert torch.equal(s1["image"], d["image"][4:]) and s1["path"] == d["path"][4:]
    with pytest.raises(AssertionError):
        tr.shard_batch(d, rank=0, world=3)
