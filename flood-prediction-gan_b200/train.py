"""Training entry point with the reference's flag surface (reference: train.py:6-37): flags map one-to-one onto
`Model(**kwargs)`. Extra flags (all optional): --batch_size / --synthetic_steps drive the synthetic loader when no
dataset is on disk; launch with torchrun for data-parallel training (one process per GPU, NCCL)."""
import argparse
import os

import torch
import torch.distributed as dist

from models import model
from models.data import SyntheticLoader

if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="Train Pix2Pix, CycleGAN, AttentionGAN or PairedAttention on the flood "
                                             "images dataset (B200-native implementation)")
    ap.add_argument("--model", required=True)
    ap.add_argument("--dataset_subset", required=True)
    ap.add_argument("--dataset_dem", required=True)
    ap.add_argument("--data_path", required=True)
    ap.add_argument("--num_epochs", type=int, default=1)
    ap.add_argument("--topography", default=None)
    ap.add_argument("--resize", type=int, default=None)
    ap.add_argument("--crop", type=int, default=None)
    ap.add_argument("--save_model_interval", type=int, default=0)
    ap.add_argument("--save_images_interval", type=int, default=0)
    ap.add_argument("--verbose", default=False, action="store_true")
    ap.add_argument("--load_pretrained_model", default=False, action="store_true")
    ap.add_argument("--pretrained_model_path", default=None)
    ap.add_argument("--add_identity_loss", action="store_true", default=False)
    ap.add_argument("--seed", type=int, default=47)
    ap.add_argument("--batch_size", type=int, default=16, help="per-GPU batch of the synthetic loader")
    ap.add_argument("--synthetic_steps", type=int, default=0, help="use a synthetic loader with this many steps/epoch")
    args = ap.parse_args()
    args.model = args.model.lower()
    if args.load_pretrained_model:
        if not args.pretrained_model_path:
            raise ValueError("Provide a saved model.")
        if not os.path.isfile(args.pretrained_model_path):
            raise FileNotFoundError("Saved model not found. Check the path to the model.")

    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl")
    kwargs = {k: v for k, v in vars(args).items() if k not in ("batch_size", "synthetic_steps")}
    kwargs["training_model"] = True
    train_model = model.Model(**kwargs)
    if args.synthetic_steps:
        size = (args.resize or 1024) // (2 if args.crop == 4 else 1)
        train_model.train_loader = SyntheticLoader(args.synthetic_steps, args.batch_size,
                                                   model.TOPOGRAPHY_CHANNELS[train_model.topography], size, rank, world)
    if train_model.model_is_cycle:
        train_model.train_cycle()
    else:
        train_model.train_paired()
    if world > 1:
        dist.destroy_process_group()
