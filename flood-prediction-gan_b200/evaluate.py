"""Evaluation entry point with the reference's flag surface (reference: evaluate.py:6-64). The generator forward and
the flood-mask thresholding run on the native kernels; plotting and torchmetrics-based image-quality metrics are out
of scope of this repository and raise NotImplementedError when requested."""
import argparse
import os

from models import model

if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="Evaluate a trained model (B200-native implementation)")
    ap.add_argument("--model", required=True)
    ap.add_argument("--dataset_subset", default="all")
    ap.add_argument("--dataset_dem", required=True)
    ap.add_argument("--use_test_data", action="store_true", default=False)
    ap.add_argument("--data_path", required=True)
    ap.add_argument("--resize", type=int, default=None)
    ap.add_argument("--crop", type=int, default=None)
    ap.add_argument("--crop_index", type=int, default=0)
    ap.add_argument("--topography", default=None)
    ap.add_argument("--pretrained_model_path", required=True)
    ap.add_argument("--plot_losses", action="store_true", default=False)
    ap.add_argument("--plot_sample_images", action="store_true", default=False)
    ap.add_argument("--num_images", type=int, default=5)
    ap.add_argument("--seed", type=int, default=47)
    ap.add_argument("--image_name", default=None)
    ap.add_argument("--plot_single_image", default=None)
    ap.add_argument("--plot_image_set", action="store_true", default=False)
    ap.add_argument("--calculate_metrics", action="store_true", default=False)
    ap.add_argument("--segmentation_model_path", default=None)
    args = ap.parse_args()
    args.model = args.model.lower()
    if not os.path.isfile(args.pretrained_model_path):
        raise FileNotFoundError("Saved model not found. Check the path to the model.")
    evaluate_model = model.Model(model=args.model, dataset_subset=args.dataset_subset, dataset_dem=args.dataset_dem,
                                 data_path=args.data_path, resize=args.resize, crop=args.crop,
                                 load_pretrained_model=True, pretrained_model_path=args.pretrained_model_path,
                                 training_model=False, seed=args.seed, topography=args.topography, verbose=True)
    if args.plot_losses or args.plot_sample_images or args.plot_single_image or args.plot_image_set:
        raise NotImplementedError("plotting is outside the accelerated hot path (DESIGN.md section 8)")
    if args.calculate_metrics:
        # flood metrics natively (U-Net inference, bit-exact masks, confusion counts); the torchmetrics image-quality
        # columns are NaN (un-vendored dependency)
        evaluate_model.calculate_metrics(use_test_data=args.use_test_data, seg_model_path=args.segmentation_model_path)
    print(f"loaded {evaluate_model.prettify_model_name()} (epoch {evaluate_model.current_epoch - 1}); "
          f"generator on {evaluate_model.device}")
