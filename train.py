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
from cocosnet_b200.util import IterationCounter, print_current_errors


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    argv = sys.argv[1:]
    if world > 1:
        argv = argv + ["--gpu_ids", str(local_rank)]
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    opt = TrainOptions().parse(argv, save=(rank == 0))
    if world == 1 and len(opt.gpu_ids) > 1:
        raise SystemExit("train.py: --gpu_ids %s names %d GPUs but this is ONE process.  Data parallelism here is one "
                         "process per GPU: launch with `torchrun --nnodes 1 --nproc-per-node %d --master-addr 127.0.0.1 "
                         "train.py ...` (the reference's nn.DataParallel is not used)."
                         % (",".join(map(str, opt.gpu_ids)), len(opt.gpu_ids), len(opt.gpu_ids)))
    assert opt.batchSize % world == 0, "batchSize must be a multiple of the number of processes"
    dataloader = cdata.create_dataloader(opt)  # same seeded shuffle on every rank: one global batch, sharded per rank
    trainer = Pix2PixTrainer(opt)
    counter = IterationCounter(opt, len(dataloader) * opt.batchSize)
    for epoch in counter.training_epochs():  # train.py:49-113
        opt.epoch = epoch
        counter.record_epoch_start(epoch)
        for i, data_i in enumerate(dataloader):
            counter.record_one_iteration()
            p = min(float(i + (epoch - 1) * len(dataloader)) / 50 / len(dataloader), 1)
            alpha = 2.0 / (1.0 + np.exp(-10 * p)) - 1
            if opt.D_steps_per_G == 1 and not (opt.weight_domainC > 0 and opt.domain_rela):
                trainer.run_step(data_i, alpha=alpha)  # CUDA-graph replay once captured (alpha is part of its key)
            else:
                if i % opt.D_steps_per_G == 0:
                    trainer.run_generator_one_step(data_i, alpha=alpha)
                trainer.run_discriminator_one_step(data_i)
            if counter.needs_printing() and rank == 0:
                print_current_errors(opt, epoch, counter.epoch_iter, trainer.get_latest_losses(), 0.0)
            if counter.needs_saving():
                trainer.save("latest")
                if rank == 0:
                    counter.record_current_iter()
        trainer.update_learning_rate(epoch)
        if rank == 0:
            counter.record_epoch_end()
        if epoch % opt.save_epoch_freq == 0 or epoch == counter.total_epochs:
            trainer.save("latest")
            trainer.save(epoch)
    print("Training was successfully finished.")
    if world > 1:
        trainer.free_graph()  # a captured NCCL all-reduce must be gone before the communicator is destroyed
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
