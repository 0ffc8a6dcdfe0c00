"""Pix2PixTrainer with the reference's methods (trainers/pix2pix_trainer.py:13-138).

Data parallelism is re-designed for one process per GPU: every rank holds a
full replica (same seed -> bit-identical init), takes its contiguous slice of
the batch (what DataParallel.scatter does) and the gradients are summed with
ONE bucketed NCCL all-reduce per optimiser step (G+Corr grads on the G step, D
grads on the D step) and divided by the world size -- the reference's
`sum(losses).mean()` over replicas.  No per-step parameter broadcast.
"""
import os

import torch
import torch.distributed as dist

from . import util
from .nets import EMA
from .pix2pix_model import Pix2PixModel

_BUCKET_BYTES = 64 << 20


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_batch(data, rank=None, world=None):
    """Contiguous equal chunk of every batched entry (DataParallel.scatter semantics)."""
    world = _world() if world is None else world
    rank = _rank() if rank is None else rank
    if world == 1:
        return data
    out = {}
    for k, v in data.items():
        n = len(v)
        assert n % world == 0, "batch %d not divisible by world size %d" % (n, world)
        per = n // world
        out[k] = v[rank * per:(rank + 1) * per]
    return out


def allreduce_grads(params, world=None):
    """Sum-then-average gradients across ranks, bucketed and flattened."""
    world = _world() if world is None else world
    if world == 1:
        return 0
    grads = [p.grad for p in params if p.grad is not None]
    buckets, cur, size = [], [], 0
    for g in grads:
        cur.append(g)
        size += g.numel() * g.element_size()
        if size >= _BUCKET_BYTES:
            buckets.append(cur)
            cur, size = [], 0
    if cur:
        buckets.append(cur)
    works = []
    for b in buckets:
        flat = torch.cat([g.reshape(-1) for g in b])
        works.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True), flat, b))
    for work, flat, b in works:
        work.wait()
        flat.div_(world)
        off = 0
        for g in b:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n
    return len(buckets)


class Pix2PixTrainer:
    def __init__(self, opt, resume_epoch=0):
        self.opt = opt
        self.pix2pix_model = Pix2PixModel(opt)
        if len(opt.gpu_ids) > 0:
            self.pix2pix_model.cuda()
            # fixed shapes every step: let cuDNN pick its fastest algorithms (measured -5 % step time on B200)
            torch.backends.cudnn.benchmark = not getattr(opt, "no_cudnn_benchmark", False)
            if getattr(opt, "channels_last", False):  # NHWC activations: no cuDNN layout transposes
                self.pix2pix_model.to(memory_format=torch.channels_last)
        self.pix2pix_model_on_one_gpu = self.pix2pix_model
        net = self.pix2pix_model.net
        if _world() > 1:  # replicas must start identical: rank 0's weights win
            for t in list(self.pix2pix_model.parameters()) + list(self.pix2pix_model.buffers()):
                dist.broadcast(t.data, src=0)
        if opt.use_ema:
            self.netG_ema, self.netCorr_ema = EMA(opt.ema_beta), EMA(opt.ema_beta)
            for ema, key in ((self.netG_ema, "netG"), (self.netCorr_ema, "netCorr")):
                for name, p in net[key].named_parameters():
                    if p.requires_grad:
                        ema.register(name, p.data)
        self.generated = None
        if opt.isTrain:
            self.optimizer_G, self.optimizer_D = self.pix2pix_model.create_optimizers(opt)
            self.old_lr = opt.lr
            if opt.continue_train and opt.which_epoch == "latest":
                ckpt = torch.load(os.path.join(opt.checkpoints_dir, opt.name, "optimizer.pth"), map_location="cpu")
                self.optimizer_G.load_state_dict(ckpt["G"])
                self.optimizer_D.load_state_dict(ckpt["D"])
            self._g_params = [p for k in ("netG", "netCorr") for p in net[k].parameters()]
            self._d_params = [p for p in net["netD"].parameters()]
        self.g_losses, self.d_losses, self.out = {}, {}, {}

    def run_generator_one_step(self, data, alpha=1):
        self.optimizer_G.zero_grad(set_to_none=True)
        # The G step only needs the gradient THROUGH the discriminator, not its weight gradients (the reference
        # computes and then discards them: optimizer_D.zero_grad() runs before they are ever used).
        for p in self._d_params:
            p.requires_grad_(False)
        try:
            g_losses, out = self.pix2pix_model(shard_batch(data), mode="generator", alpha=alpha)
            g_loss = sum(g_losses.values()).mean()
            g_loss.backward()
        finally:
            for p in self._d_params:
                p.requires_grad_(True)
        allreduce_grads(self._g_params)
        self.optimizer_G.step()
        self.g_losses, self.out = g_losses, out
        if self.opt.use_ema:
            self.netG_ema(self.pix2pix_model.net["netG"])
            self.netCorr_ema(self.pix2pix_model.net["netCorr"])

    def run_discriminator_one_step(self, data):
        self.optimizer_D.zero_grad(set_to_none=True)
        GforD = {k: self.out.get(k) for k in ("fake_image", "adaptive_feature_seg", "adaptive_feature_img")}
        d_losses = self.pix2pix_model(shard_batch(data), mode="discriminator", GforD=GforD)
        d_loss = sum(d_losses.values()).mean()
        d_loss.backward()
        allreduce_grads(self._d_params)
        self.optimizer_D.step()
        self.d_losses = d_losses

    def get_latest_losses(self):
        return {**self.g_losses, **self.d_losses}

    def get_latest_generated(self):
        return self.out["fake_image"]

    def save(self, epoch):
        if _rank() != 0:
            return
        model = self.pix2pix_model
        model.save(epoch)
        if self.opt.use_ema:
            for ema, key, label in ((self.netG_ema, "netG", "G_ema"), (self.netCorr_ema, "netCorr", "netCorr_ema")):
                ema.assign(model.net[key])
                util.save_network(model.net[key], label, epoch, self.opt)
                ema.resume(model.net[key])
        if epoch == "latest":
            torch.save({"G": self.optimizer_G.state_dict(), "D": self.optimizer_D.state_dict(), "lr": self.old_lr},
                       os.path.join(self.opt.checkpoints_dir, self.opt.name, "optimizer.pth"))

    def update_learning_rate(self, epoch):
        new_lr = self.old_lr - self.opt.lr / self.opt.niter_decay if epoch > self.opt.niter else self.old_lr
        if new_lr != self.old_lr:
            g_lr, d_lr = (new_lr, new_lr) if self.opt.no_TTUR else (new_lr / 2, new_lr * 2)
            for pg in self.optimizer_D.param_groups:
                pg["lr"] = d_lr
            for pg in self.optimizer_G.param_groups:
                pg["lr"] = g_lr
            print("update learning rate: %f -> %f" % (self.old_lr, new_lr))
            self.old_lr = new_lr
