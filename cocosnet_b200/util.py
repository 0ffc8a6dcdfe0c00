"""Small helpers with the reference's semantics (util/util.py)."""
import importlib
import os
import sys

import torch

_EPS = sys.float_info.epsilon


def feature_normalize(x):
    """x / (||x||_2 over dim 1 + eps)   (util/util.py:31-34)."""
    return x / (torch.norm(x, 2, 1, keepdim=True) + _EPS)


def weighted_l1_loss(inp, target, weights):  # util/util.py:36-40
    return (torch.abs(inp - target) * weights.expand_as(inp)).mean()


def mse_loss(inp, target=0):  # util/util.py:42-43
    return torch.mean((inp - target) ** 2)


_VGG_MEAN = (0.40760392, 0.45795686, 0.48501961)


_VGG_MEAN_CACHE = {}


def vgg_preprocess(t, vgg_normal_correct=False):
    """RGB in [0,1] (or [-1,1] with vgg_normal_correct) -> BGR, mean-subtracted,
    x255 (util/util.py:45-54)."""
    if vgg_normal_correct:
        t = (t + 1) / 2
    bgr = t.flip(1)
    key = (t.dtype, t.device)
    mean = _VGG_MEAN_CACHE.get(key)
    if mean is None:  # built once per device: a host->device copy of a Python list cannot be captured in a CUDA graph
        mean = _VGG_MEAN_CACHE[key] = torch.tensor(_VGG_MEAN, dtype=t.dtype, device=t.device).view(1, 3, 1, 1)
    return (bgr - mean) * 255


def find_class_in_module(target_cls_name, module):
    """Case-insensitive class lookup by name with '_' removed (util/util.py:211-223)."""
    target = target_cls_name.replace("_", "").lower()
    mod = importlib.import_module(module) if isinstance(module, str) else module
    for name, cls in mod.__dict__.items():
        if name.lower() == target and isinstance(cls, type):
            return cls
    raise ValueError("In %s there should be a class whose lower-case name is %s" % (module, target))


def save_network(net, label, epoch, opt):
    """'<epoch>_net_<label>.pth' plain state_dict (util/util.py:226-231)."""
    path = os.path.join(opt.checkpoints_dir, opt.name, "%s_net_%s.pth" % (epoch, label))
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save({k: v.detach().cpu() for k, v in net.state_dict().items()}, path)


def load_network(net, label, epoch, opt):
    """Tolerant load (util/util.py:234-250): missing file -> untouched net,
    shape/key mismatch -> strict=False."""
    path = os.path.join(opt.checkpoints_dir, opt.name, "%s_net_%s.pth" % (epoch, label))
    if not os.path.exists(path):
        print("not find model :" + path + ", do not load model!")
        return net
    weights = torch.load(path, map_location="cpu")
    try:
        net.load_state_dict(weights)
    except KeyError:
        print("key error, not load!")
    except RuntimeError as err:
        print(err)
        net.load_state_dict(weights, strict=False)
        print("loaded with strict=False")
    return net


def print_current_errors(opt, epoch, i, errors, t):  # util/util.py:320-331
    msg = "(epoch: %d, iters: %d, time: %.3f) " % (epoch, i, t)
    for k, v in errors.items():
        msg += "%s: %.3f " % (k, v.mean().float())
    print(msg)
    try:
        with open(os.path.join(opt.checkpoints_dir, opt.name, "loss_log.txt"), "a") as f:
            f.write("%s\n" % msg)
    except OSError as err:
        print(err)


class IterationCounter:
    """Epoch / iteration bookkeeping with resume (util/iter_counter.py:12-74): `<checkpoints>/<name>/iter.txt` holds
    "epoch,iter_in_epoch"; --continue_train restarts from it, so the linear lr decay, `alpha` and the mask_epoch switch
    pick up mid-schedule instead of at epoch 1."""

    def __init__(self, opt, dataset_size):
        import numpy as np
        self.opt, self.dataset_size = opt, dataset_size
        self.first_epoch, self.epoch_iter = 1, 0
        self.total_epochs = opt.niter + opt.niter_decay
        self.iter_record_path = os.path.join(opt.checkpoints_dir, opt.name, "iter.txt")
        if opt.isTrain and opt.continue_train:
            try:
                self.first_epoch, self.epoch_iter = (int(v) for v in np.loadtxt(self.iter_record_path, delimiter=",",
                                                                                 dtype=int))
                print("Resuming from epoch %d at iteration %d" % (self.first_epoch, self.epoch_iter))
            except (OSError, ValueError):
                print("Could not load iteration record at %s. Starting from beginning." % self.iter_record_path)
        self.total_steps_so_far = (self.first_epoch - 1) * dataset_size + self.epoch_iter
        self.current_epoch = self.first_epoch

    def training_epochs(self):
        return range(self.first_epoch, self.total_epochs + 1)

    def record_epoch_start(self, epoch):
        self.epoch_iter, self.current_epoch = 0, epoch

    def record_one_iteration(self):
        self.total_steps_so_far += self.opt.batchSize
        self.epoch_iter += self.opt.batchSize

    def _write(self, epoch, it):
        os.makedirs(os.path.dirname(self.iter_record_path), exist_ok=True)
        with open(self.iter_record_path, "w") as f:
            f.write("%d\n%d\n" % (epoch, it))  # np.savetxt's layout: one value per line

    def record_epoch_end(self):
        if self.current_epoch % self.opt.save_epoch_freq == 0:
            self._write(self.current_epoch + 1, 0)

    def record_current_iter(self):
        self._write(self.current_epoch, self.epoch_iter)

    def needs_saving(self):
        return (self.total_steps_so_far % self.opt.save_latest_freq) < self.opt.batchSize

    def needs_printing(self):
        return (self.total_steps_so_far % self.opt.print_freq) < self.opt.batchSize
