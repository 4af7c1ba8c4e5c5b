"""Evaluation entry point with the reference's flag surface (reference: evaluate.py:6-64). --calculate_metrics runs
generator inference, the segmentation U-Net, the bit-exact flood masks / confusion metrics and PSNR / SSIM / MS-SSIM on
the native kernels (LPIPS is NaN: it needs downloaded AlexNet weights); the plotting flags are accepted but plotting is
out of scope of this repository and raises NotImplementedError."""
import os

from models import model
from train import REQUIRED, build_parser

FLAGS = (("model", str, REQUIRED), ("dataset_subset", str, "all"), ("dataset_dem", str, REQUIRED),
         ("use_test_data", None, False), ("data_path", str, REQUIRED), ("resize", int, None), ("crop", int, None),
         ("crop_index", int, 0), ("topography", str, None), ("pretrained_model_path", str, REQUIRED),
         ("plot_losses", None, False), ("plot_sample_images", None, False), ("num_images", int, 5), ("seed", int, 47),
         ("image_name", str, None), ("plot_single_image", str, None), ("plot_image_set", None, False),
         ("calculate_metrics", None, False), ("segmentation_model_path", str, None))


def main():
    args = build_parser("B200-native evaluation of a trained model", FLAGS).parse_args()
    if not os.path.isfile(args.pretrained_model_path):
        raise FileNotFoundError("Saved model not found. Check the path to the model.")
    net = model.Model(model=args.model.lower(), dataset_subset=args.dataset_subset, dataset_dem=args.dataset_dem,
                      data_path=args.data_path, resize=args.resize, crop=args.crop, load_pretrained_model=True,
                      pretrained_model_path=args.pretrained_model_path, training_model=False, seed=args.seed,
                      topography=args.topography, verbose=True)
    if args.plot_losses or args.plot_sample_images or args.plot_single_image or args.plot_image_set:
        raise NotImplementedError("plotting is outside the accelerated hot path (DESIGN.md section 8)")
    if args.calculate_metrics:
        net.calculate_metrics(use_test_data=args.use_test_data, seg_model_path=args.segmentation_model_path)
    print(f"loaded {net.prettify_model_name()} (epoch {net.current_epoch - 1}); generator on {net.device}")


if __name__ == "__main__":
    main()
