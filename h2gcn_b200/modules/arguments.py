"""Hookable argparse — mirrors `h2gcn/modules/arguments.py` (create_parser / parse_args, the `function_hooks["argparse"]`
deque run after parsing, and `args.objects` as the service registry).  signac integration is out of scope."""
import argparse
from collections import deque


def create_parser():
    parser = argparse.ArgumentParser(add_help=False)
    parser.function_hooks = dict()
    parser.function_hooks["argparse"] = deque()
    return parser


def parse_args(parser, argv=None):
    parser.add_argument("--verbose", "-v", action="store_true")
    parser.add_argument("--help", "-h", action="help")
    parser.add_argument("--exp_tags", default=[], nargs="+", dest="_exp_tags")
    args = parser.parse_args(argv)
    args.use_signac = False
    args.objects = dict(function_hooks=parser.function_hooks)
    for key in ("pretrain_callbacks", "pre_epoch_callbacks", "post_epoch_callbacks", "post_train_callbacks"):
        args.objects[key] = deque()
    while len(parser.function_hooks["argparse"]) > 0:   # dataset hook first (appendleft), then the model hook
        parser.function_hooks["argparse"].popleft()(args)
    return args
