"""Training entry point with the reference's flag surface (reference: train.py:6-37): flags map one-to-one onto
`Model(**kwargs)`. Extra flags (all optional): --batch_size (per-GPU batch of the device-resident loaders, default 1 as
in the reference) and --synthetic_steps (synthetic loader when no dataset is on disk); launch with torchrun for
data-parallel training (one process per GPU, NCCL)."""
import argparse
import os

import torch
import torch.distributed as dist

from models import model
from models.data import SyntheticLoader

# (flag, type or None for a switch, default or REQUIRED) -- the reference's flags in its order, then the extras
REQUIRED = object()
FLAGS = (("model", str, REQUIRED), ("dataset_subset", str, REQUIRED), ("dataset_dem", str, REQUIRED),
         ("data_path", str, REQUIRED), ("num_epochs", int, 1), ("topography", str, None), ("resize", int, None),
         ("crop", int, None), ("save_model_interval", int, 0), ("save_images_interval", int, 0), ("verbose", None, False),
         ("load_pretrained_model", None, False), ("pretrained_model_path", str, None), ("add_identity_loss", None, False),
         ("seed", int, 47), ("batch_size", int, 1), ("synthetic_steps", int, 0))


def build_parser(description, flags):
    parser = argparse.ArgumentParser(description=description)
    for name, kind, default in flags:
        if kind is None:
            parser.add_argument("--" + name, action="store_true", default=default)
        elif default is REQUIRED:
            parser.add_argument("--" + name, type=kind, required=True)
        else:
            parser.add_argument("--" + name, type=kind, default=default)
    return parser


def main():
    args = build_parser("B200-native training of Pix2Pix / CycleGAN / AttentionGAN / PairedAttention", FLAGS).parse_args()
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
    kwargs = {k: v for k, v in vars(args).items() if k != "synthetic_steps"}
    net = model.Model(training_model=True, **kwargs)
    if args.synthetic_steps:
        size = (args.resize or 1024) // (2 if args.crop == 4 else 1)
        net.train_loader = SyntheticLoader(args.synthetic_steps, args.batch_size,
                                           model.TOPOGRAPHY_CHANNELS[net.topography], size, rank, world)
    if args.save_images_interval and rank == 0:
        print("note: --save_images_interval is accepted for flag compatibility; plotting is outside this build")
    try:
        (net.train_cycle if net.model_is_cycle else net.train_paired)()
    finally:
        net.close()  # captured NCCL work must be released before the process group goes away
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
