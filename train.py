#!/usr/bin/env python
"""Training entry point with the reference's CLI (reference train.py:17-122).
Single GPU:  python train.py --name ade20k --dataset_mode ade20k --use_attention --maskmix --PONO --PONO_C ...
8 GPUs:      torchrun --nnodes 1 --nproc-per-node 8 --master-addr 127.0.0.1 train.py ... --batchSize 64
(one process per GPU; --gpu_ids is replaced by LOCAL_RANK, --batchSize is the global batch as in the reference)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

from cocosnet_b200 import data as cdata
from cocosnet_b200.options import TrainOptions
from cocosnet_b200.trainer import Pix2PixTrainer
from cocosnet_b200.util import print_current_errors


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    argv = sys.argv[1:]
    if world > 1:
        argv = argv + ["--gpu_ids", str(local_rank)]
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    opt = TrainOptions().parse(argv, save=(int(os.environ.get("RANK", "0")) == 0))
    assert opt.batchSize % world == 0, "batchSize must be a multiple of the number of processes"
    dataloader = cdata.create_dataloader(opt)
    trainer = Pix2PixTrainer(opt)
    total_epochs = opt.niter + opt.niter_decay
    steps = 0
    for epoch in range(1, total_epochs + 1):
        opt.epoch = epoch
        for i, data_i in enumerate(dataloader):
            p = min(float(i + (epoch - 1) * len(dataloader)) / 50 / len(dataloader), 1)
            alpha = 2.0 / (1.0 + np.exp(-10 * p)) - 1
            if opt.D_steps_per_G == 1 and not (opt.weight_domainC > 0 and opt.domain_rela):
                trainer.run_step(data_i, alpha=alpha)  # alpha unused here; one GPU: CUDA-graph replay
            else:
                if i % opt.D_steps_per_G == 0:
                    trainer.run_generator_one_step(data_i, alpha=alpha)
                trainer.run_discriminator_one_step(data_i)
            steps += opt.batchSize
            if steps % opt.print_freq < opt.batchSize and int(os.environ.get("RANK", "0")) == 0:
                print_current_errors(opt, epoch, i, trainer.get_latest_losses(), 0.0)
            if steps % opt.save_latest_freq < opt.batchSize:
                trainer.save("latest")
        trainer.update_learning_rate(epoch)
        if epoch % opt.save_epoch_freq == 0 or epoch == total_epochs:
            trainer.save("latest")
            trainer.save(epoch)
    print("Training was successfully finished.")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
