"""CLI — mirrors the argv surface of `h2gcn/run_experiments.py`:

    python -m h2gcn_b200.run_experiments H2GCN planetoid --dataset ind.cora --dataset_path <dir> \
        [--network_setup M64-R-T1-G-V-T2-G-V-C1-C2-D0.5-MO --hidden 64 --adj_nhood 1 2 --no_feature_normalize ...]

Builds the parser from the model / dataset plug-ins, runs the argparse hooks (dataset load -> preprocessing on the GPU
-> model construction) and ONE forward pass (`predict_step` + `test_step`).  The epoch loop of the reference
(`run_experiments.py:47-61`: optimizer, checkpoints, early stopping) is outside the accelerated path."""
import sys
import time


def main(argv=None):
    import torch
    from h2gcn_b200 import datasets, models
    from h2gcn_b200.modules import arguments
    parser = arguments.create_parser()
    parser.add_argument("--random_seed", type=int, default=123)
    parser.add_argument("--epochs", type=int, default=2000, help="accepted for compatibility; no training loop here")
    old_argv = sys.argv
    if argv is not None:
        sys.argv = [old_argv[0]] + list(argv)
    try:
        models.add_subparsers(parser)
        datasets.add_subparsers(parser)
        args = arguments.parse_args(parser)
    finally:
        sys.argv = old_argv
    if args.random_seed:
        torch.manual_seed(args.random_seed)
    tensors = args.objects["tensors"]
    t0 = time.perf_counter()
    logits = args.objects["predict_step"](**tensors)
    torch.cuda.synchronize()
    stats = args.objects["test_step"](**tensors)
    print(f"forward: logits {tuple(logits.shape)} in {1e3 * (time.perf_counter() - t0):.2f} ms (first call, includes planning); "
          f"random-weight accuracies {stats}")
    return args, logits


if __name__ == "__main__":
    main()
