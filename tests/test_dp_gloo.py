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


# ---- batch-statistics SPADE across ranks (the reference's SynchronizedBatchNorm2d, normalization.py:96-104) ----
SYNCBN_ARGV = ["--dataset_mode", "celebahq", "--warp_bilinear", "--adaptor_kernel", "4", "--gpu_ids", "-1",
               "--crop_size", "64", "--load_size", "64", "--batchSize", "2"]


def _celebahq_grads(shard=None):
    """One generator pass + backward of the celebahq model (SPADE with batch statistics, no --PONO) at 64x64 with the
    tape on the kernel emulation; shard = (rank, world): this rank's slice of the 2-image batch, gradients averaged
    over the ranks as the trainer does.  Returns (losses, gradients, BatchNorm running estimates)."""
    from cocosnet_b200 import data as cdata
    from cocosnet_b200 import nhwc
    from cocosnet_b200 import trainer as tr
    from cocosnet_b200.options import TrainOptions
    from cocosnet_b200.pix2pix_model import Pix2PixModel
    from oracle import torch_port
    from oracle.nhwc_emul import EmulBackend
    old = nhwc.set_backend(EmulBackend(exact=True))
    try:
        opt = TrainOptions().parse(SYNCBN_ARGV, save=False, verbose=False)
        opt.verbose_networks = False
        opt.allow_random_vgg = True
        torch.manual_seed(0)
        model = Pix2PixModel(opt)
        model.vggnet_fix.load_state_dict(cdata.seeded_vgg_state_dict())
        model.train()
        batch = cdata.synthetic_batch(opt, 2)
        if shard is not None:
            batch = tr.shard_batch(batch, rank=shard[0], world=shard[1])
        with torch_port.cpu_reference_mode():
            g_losses, _ = model(batch, mode="generator")
            sum(g_losses.values()).mean().backward()
        params = [p for k in ("netG", "netCorr") for p in model.net[k].parameters() if p.grad is not None]
        if shard is not None:
            tr.allreduce_grads(params)
        grads = {k + "/" + n: p.grad.clone() for k in ("netG", "netCorr") for n, p in model.net[k].named_parameters()
                 if p.grad is not None}
        stats = {k + "/" + n: b.clone() for k in ("netG", "netCorr") for n, b in model.net[k].named_buffers()
                 if n.endswith(("running_mean", "running_var"))}
        return {k: float(v.detach().mean()) for k, v in g_losses.items()}, grads, stats
    finally:
        nhwc.set_backend(old)


def _syncbn_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(4)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.save(_celebahq_grads((rank, world)), os.path.join(out_dir, "syncbn%d.pt" % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_batch_statistics_spade_is_synchronised_over_ranks(tmp_path):
    """Two ranks with one image each == one process with both images: the SPADE layers' batch statistics (forward) and
    the two reductions of their backward are averaged over the ranks inside the tape (tape.spade_stat), so losses,
    averaged gradients and the running estimates agree with the global-batch computation.  Unsynchronised statistics
    (one image instead of two) would move all three by tens of percent."""
    port = _free_port()
    mp.spawn(_syncbn_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    (l0, g0, s0), (l1, g1, s1) = (torch.load(os.path.join(tmp_path, "syncbn%d.pt" % r)) for r in (0, 1))
    lw, gw, sw = _celebahq_grads()
    assert len(sw) > 40
    for k in sw:  # same estimates on both ranks, equal to the global-batch ones
        assert torch.equal(s0[k], s1[k]), k
        assert torch.allclose(s0[k], sw[k], rtol=1e-4, atol=1e-6), (k, float((s0[k] - sw[k]).abs().max()))
    for k in lw:
        assert abs(0.5 * (l0[k] + l1[k]) - lw[k]) <= 2e-4 * abs(lw[k]) + 1e-6, (k, l0[k], l1[k], lw[k])
    med = sorted(float(v.norm()) for v in gw.values())[len(gw) // 2]
    checked = 0
    for k, v in gw.items():
        assert torch.equal(g0[k], g1[k]), k
        if float(v.norm()) < 1e-2 * med:
            continue  # analytically zero (a bias in front of a normalisation): rounding noise
        d = float((g0[k] - v).norm() / v.norm())
        assert d < 2e-2, (k, d)
        checked += 1
    assert checked > 150


# ---- the whole trainer iteration at world size 2 (the tape on the kernel emulation) ----
TRAINER_ARGV = ["--dataset_mode", "ade20k", "--PONO", "--PONO_C", "--use_attention", "--maskmix", "--gpu_ids", "-1",
                "--crop_size", "64", "--load_size", "64", "--batchSize", "2"]


def _trainer_grads():
    """run_generator_one_step + run_discriminator_one_step of the real trainer (shard, backward, all-reduce) with the
    optimiser steps replaced by recorders: the gradients each optimiser would have consumed."""
    from cocosnet_b200 import data as cdata
    from cocosnet_b200 import nhwc
    from cocosnet_b200.options import TrainOptions
    from cocosnet_b200.trainer import Pix2PixTrainer
    from oracle import torch_port
    from oracle.nhwc_emul import EmulBackend
    old = nhwc.set_backend(EmulBackend(exact=True))
    try:
        opt = TrainOptions().parse(TRAINER_ARGV, save=False, verbose=False)
        opt.verbose_networks = False
        opt.allow_random_vgg = True
        torch.manual_seed(0)
        trainer = Pix2PixTrainer(opt)
        model = trainer.pix2pix_model
        model.vggnet_fix.load_state_dict(cdata.seeded_vgg_state_dict())
        seen = {}

        def recorder(nets):
            def step():
                for k in nets:
                    for n, p in model.net[k].named_parameters():
                        if p.grad is not None:
                            seen[k + "/" + n] = p.grad.clone()
            return step
        trainer.optimizer_G.step = recorder(("netG", "netCorr"))
        trainer.optimizer_D.step = recorder(("netD",))
        batch = cdata.synthetic_batch(opt, 2)  # the GLOBAL batch on every rank: the trainer takes this rank's slice
        with torch_port.cpu_reference_mode():
            trainer.run_generator_one_step(batch)
            trainer.run_discriminator_one_step(batch)
        losses = {k: float(v.detach().mean()) for k, v in trainer.get_latest_losses().items()}
        return losses, seen
    finally:
        nhwc.set_backend(old)


def _trainer_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(4)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.save(_trainer_grads(), os.path.join(out_dir, "trainer%d.pt" % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_trainer_iteration_on_two_ranks_matches_the_global_batch(tmp_path):
    """The reference's nn.DataParallel step (pix2pix_trainer.py:23-26,52-74) as one process per GPU: each rank runs the
    G and D steps on its contiguous half of the batch and averages the gradients.  Every loss term of this flag set is a
    mean over samples, so the averaged gradients equal those of one process on the whole batch."""
    port = _free_port()
    mp.spawn(_trainer_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    (l0, g0), (l1, g1) = (torch.load(os.path.join(tmp_path, "trainer%d.pt" % r)) for r in (0, 1))
    lw, gw = _trainer_grads()
    assert set(g0) == set(gw) and len(gw) > 300
    for k in lw:
        assert abs(0.5 * (l0[k] + l1[k]) - lw[k]) <= 2e-4 * abs(lw[k]) + 1e-6, (k, l0[k], l1[k], lw[k])
    med = sorted(float(v.norm()) for v in gw.values())[len(gw) // 2]
    checked = 0
    for k, v in gw.items():
        assert torch.equal(g0[k], g1[k]), k
        if float(v.norm()) < 1e-2 * med:
            continue
        d = float((g0[k] - v).norm() / v.norm())
        assert d < 2e-2, (k, d)
        checked += 1
    assert checked > 250
