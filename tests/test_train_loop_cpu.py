"""CPU: the pieces of train.py around the step -- resume bookkeeping (util/iter_counter.py:12-74), the rank-consistent
loader, the graph key of Pix2PixTrainer.run_step, EMA in place."""
import os

import torch

from cocosnet_b200 import data as cdata
from cocosnet_b200.options import TrainOptions
from cocosnet_b200.util import IterationCounter

ARGV = ["--dataset_mode", "ade20k", "--use_attention", "--maskmix", "--PONO", "--PONO_C", "--batchSize", "2", "--gpu_ids", "-1"]


def _opt(tmp, extra=()):
    opt = TrainOptions().parse(ARGV + ["--checkpoints_dir", str(tmp), "--name", "t"] + list(extra), save=False, verbose=False)
    return opt


def test_iteration_counter_resumes_mid_schedule(tmp_path):
    opt = _opt(tmp_path)
    c = IterationCounter(opt, dataset_size=64)
    assert list(c.training_epochs())[0] == 1 and c.total_steps_so_far == 0
    c.record_epoch_start(3)
    for _ in range(5):
        c.record_one_iteration()
    c.record_current_iter()
    assert open(os.path.join(tmp_path, "t", "iter.txt")).read().split() == ["3", "10"]
    opt2 = _opt(tmp_path, ["--continue_train"])
    c2 = IterationCounter(opt2, dataset_size=64)
    assert c2.first_epoch == 3 and c2.epoch_iter == 10 and c2.total_steps_so_far == 2 * 64 + 10
    assert list(c2.training_epochs())[0] == 3
    c2.record_epoch_start(10)
    c2.record_epoch_end()  # epoch 10 % save_epoch_freq (10) == 0 -> next epoch recorded
    assert open(os.path.join(tmp_path, "t", "iter.txt")).read().split() == ["11", "0"]


def test_every_rank_draws_the_same_global_batch(tmp_path):
    opt = _opt(tmp_path)
    a = [b["path"] for b in cdata.create_dataloader(opt)]
    b = [b["path"] for b in cdata.create_dataloader(opt)]
    assert a == b and len(a) > 2 and a[0] != a[1]


def test_ema_updates_in_place():
    from cocosnet_b200.nets import EMA
    lin = torch.nn.Linear(3, 2)
    ema = EMA(0.9)
    for n, p in lin.named_parameters():
        ema.register(n, p.data)
    ptrs = {n: t.data_ptr() for n, t in ema.shadow.items()}
    w0 = lin.weight.data.clone()
    with torch.no_grad():
        lin.weight.add_(1.0)
    ema(lin)
    assert {n: t.data_ptr() for n, t in ema.shadow.items()} == ptrs
    assert torch.allclose(ema.shadow["weight"], 0.9 * w0 + 0.1 * (w0 + 1))
    wptr = lin.weight.data_ptr()
    ema.assign(lin)
    assert lin.weight.data_ptr() == wptr and torch.allclose(lin.weight.data, ema.shadow["weight"])
    ema.resume(lin)
    assert lin.weight.data_ptr() == wptr and torch.allclose(lin.weight.data, w0 + 1)


def test_netG_gradients_are_final_when_the_overlap_hook_fires(monkeypatch, tmp_path):
    """The multi-rank schedule starts netG's gradient all-reduce from an autograd hook on netCorr's output.  That is only
    sound if, when the hook fires, every netG parameter already holds its FINAL gradient of the step.  Checked here on
    the CPU autograd engine (same engine, same node priorities as on the GPU) with the collectives replaced by
    recorders: the early set is exactly netG's parameters, their gradients do not change afterwards, the rest is
    reduced after the backward and the optimiser steps last."""
    from cocosnet_b200 import trainer as tmod
    from cocosnet_b200.trainer import Pix2PixTrainer
    opt = TrainOptions().parse(["--dataset_mode", "ade20k", "--use_attention", "--maskmix", "--PONO", "--PONO_C",
                                "--batchSize", "1", "--gpu_ids", "-1", "--checkpoints_dir", str(tmp_path), "--name", "t"],
                               save=False, verbose=False)
    opt.verbose_networks = False
    opt.allow_random_vgg = True
    torch.manual_seed(0)
    trainer = Pix2PixTrainer(opt)
    trainer.pre_sharded = True  # the batch below is this "rank"'s shard
    batch = cdata.synthetic_batch(opt, 1)
    events = []

    class Handle:
        def wait(self):
            events.append(("wait",))

    def fake_allreduce(params, world=None, async_op=False):
        params = list(params)
        events.append(("allreduce", async_op, [id(p) for p in params], [p.grad.clone() for p in params]))
        return Handle() if async_op else None

    monkeypatch.setattr(tmod, "allreduce_grads", fake_allreduce)
    monkeypatch.setattr(tmod, "_world", lambda: 2)
    monkeypatch.setattr(tmod.dist, "get_backend", lambda *a, **k: "nccl")
    real_step = trainer.optimizer_G.step
    monkeypatch.setattr(trainer.optimizer_G, "step", lambda *a, **k: (events.append(("step",)), real_step(*a, **k))[1])
    from oracle import torch_port
    with torch_port.cpu_reference_mode():
        trainer.run_generator_one_step(batch)
    kinds = [e[0] for e in events]
    assert kinds == ["allreduce", "allreduce", "wait", "step"], kinds
    early, late = events[0], events[1]
    netg = list(trainer.pix2pix_model.net["netG"].parameters())
    netc = list(trainer.pix2pix_model.net["netCorr"].parameters())
    assert early[1] is True and early[2] == [id(p) for p in netg if p.requires_grad]
    assert late[1] is False and late[2] == [id(p) for p in netc]
    for p, g_then in zip([p for p in netg if p.requires_grad], early[3]):
        assert torch.equal(p.grad, g_then), "a netG gradient changed after the hook fired"
