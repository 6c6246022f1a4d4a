"""Dataset package — mirrors `h2gcn/datasets/__init__.py`: positional `datafmt` argument + plug-in discovery."""
import contextlib
import importlib
import os
import pkgutil

from ._dataset import TransformSPAdj  # noqa: F401


def add_subparsers(parser):
    dataset_list = [name for _, name, _ in pkgutil.iter_modules(path=__path__) if not name.startswith("_")]
    parser.add_argument("datafmt", choices=dataset_list, help="Dataset selected for experiment")
    try:
        with open(os.devnull, "w") as null, contextlib.redirect_stderr(null):
            known, _ = parser.parse_known_args()
    except SystemExit:
        return
    module = importlib.import_module("." + known.datafmt, package=__name__)
    if hasattr(module, "add_subparser_args"):
        module.add_subparser_args(parser)
