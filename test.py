#!/usr/bin/env python
"""Inference entry point with the reference's CLI (reference test.py:15-69):
loads <checkpoints_dir>/<name>/<which_epoch>_net_{G,Corr}.pth (reference checkpoints load unchanged)
and writes label | exemplar | warp | output grids."""
import os
import sys

import torch
import torchvision.utils as vutils

from cocosnet_b200 import data as cdata
from cocosnet_b200.options import TestOptions
from cocosnet_b200.pix2pix_model import Pix2PixModel


def main():
    opt = TestOptions().parse(sys.argv[1:])
    torch.manual_seed(0)
    dataloader = cdata.create_dataloader(opt)
    model = Pix2PixModel(opt)
    if len(opt.gpu_ids) > 0:
        model.cuda()
    model.eval()
    save_root = os.path.join(os.path.dirname(opt.checkpoints_dir), "output", opt.name)
    os.makedirs(save_root, exist_ok=True)
    for i, data_i in enumerate(dataloader):
        if i * opt.batchSize >= opt.how_many:
            break
        out = model(data_i, mode="inference")
        imgs = torch.cat((data_i["ref"], out["warp_out"].cpu(), out["fake_image"].cpu()), 0)
        vutils.save_image(imgs, os.path.join(save_root, "%d.png" % i), nrow=data_i["ref"].shape[0], padding=0,
                          normalize=True)
    print("wrote", save_root)


if __name__ == "__main__":
    main()
