"""Planetoid dataset plug-in — mirrors `h2gcn/datasets/planetoid.py` (same argv names; registers itself FIRST in the
argparse hook deque so the model hook finds `args.objects["dataset"]`)."""
from ._dataset import PlanetoidData


def add_subparser_args(parser):
    sub = parser.add_argument_group("Planetoid Format Data Arguments (datasets/planetoid.py)")
    sub.add_argument("--dataset", type=str, required=True)
    sub.add_argument("--dataset_path", type=str, dest="_dataset_path", required=True)
    sub.add_argument("--val_size", type=int, default=500)
    sub.add_argument("--feature_configs", choices=["no_test"], nargs="*", default=[])
    parser.function_hooks["argparse"].appendleft(argparse_callback)


def argparse_callback(args):
    if args.val_size < 0:
        args.val_size = None
    dataset = PlanetoidData(args.dataset, args._dataset_path, val_size=args.val_size)
    for config in args.feature_configs:
        if config == "no_test":
            lil = dataset.features.tolil()
            lil[dataset.test_mask, :] = 0
            dataset.features = lil.tocsr()
    args.objects["dataset"] = dataset
    print(f"===> Dataset loaded: {args.dataset}")
