"""Model package of the B200-native H2GCN path — mirrors `h2gcn/models/__init__.py` of the reference.

Public names kept: `Layer`, `parse_network_setup`, `toNumpy`, `add_subparsers`.
`parse_network_setup` implements the grammar of the reference's layer-string DSL (h2gcn/models/__init__.py:47-150);
tests/test_parse_network_setup.py checks it against the reference's own output for every string in the reference's
experiment configs (tests/golden/parse_network_setup.json).
"""
import re


class Layer:
    """Layer type codes (reference: models/__init__.py:34-44)."""
    DENSE = "F"
    DROPOUT = "D"
    GCN = "G"
    RELU = "R"
    CONCAT = "C"
    VECTORIZE = "V"
    IDENTITY = "I"
    SLICE = "S"
    EXPERIMENTAL = "X"
    LAMBDA = "lambda"


_SPLIT = re.compile(r"-(?![^[]*\])")          # '-' outside [...]
_SLICE = re.compile(r"^S([^_]*)(?:_|$)((?:[^_]*(?:_|$))*)")
_XNAME = re.compile(r"X([^_]*)(?:_|$)(.*)")


def _dense(body, use_bias, output_dim, default_units):
    conf = {}
    if body == "O":                             # "MO"/"FO": output layer, units = number of labels (:64-66)
        units = output_dim
        conf["beginOutput"] = True
    elif body:
        units = int(body)
    else:
        assert default_units is not None
        units = default_units
    return Layer.DENSE, dict(units=units, use_bias=use_bias, **conf)


def parse_network_setup(network_setup_str, output_dim, _dense_units=None, _dropout_rate=None,
                        parse_preprocessing=False):
    """'M64-R-T1-G-V-T2-G-V-C1-C2-D0.5-MO' -> [(layerType, conf), ...]   (reference models/__init__.py:47-150).

    Tokens: F<n>/M<n> dense with/without bias ('O' = output width), D<p> dropout, G[i_j..] hop aggregation,
    C<t1_t2..> concat with tagged tensors, R relu, V flatten, I to-dense, S<tag>_<start>_<stop>_<step> slice,
    X<name>_<conf> experimental, lambda..., and the modifiers E (embedding), L (supervised), T<tag> which annotate
    the PREVIOUS layer.  Unknown tokens raise ValueError."""
    tokens = _SPLIT.split(network_setup_str)
    conf_list = []
    embedding_defined = False
    for tok in tokens:
        if tok[0] == "[" and tok[-1] == "]":
            tok = tok[1:-1].strip()
        head, body = tok[0], tok[1:]
        if tok.startswith("lambda"):
            conf_list.append((Layer.LAMBDA, {"lambda": tok}))
        elif head in ("F", "M"):
            conf_list.append(_dense(body, head == "F", output_dim, _dense_units))
        elif head == "D":
            if body:
                rate = float(body)
            else:
                assert _dropout_rate is not None
                rate = _dropout_rate
            conf_list.append((Layer.DROPOUT, dict(dropout_rate=rate)))
        elif head == "G":
            hops = set(int(i) for i in body.split("_")) if body else None
            conf_list.append((Layer.GCN, dict(hops=hops)))
        elif head == "C":
            conf_list.append((Layer.CONCAT, dict(tags=list(body.split("_")), addInputs=True)))
        elif head == "R":
            conf_list.append((Layer.RELU, dict()))
        elif head == "V":
            conf_list.append((Layer.VECTORIZE, dict()))
        elif head == "I":
            conf_list.append((Layer.IDENTITY, dict()))
        elif head == "S":
            m = _SLICE.search(tok)
            tag = m.group(1) or None
            if m.group(2):
                slice_obj = slice(*[(int(x) if x else None) for x in m.group(2).split("_")])
            else:
                slice_obj = slice(None)
            conf_list.append((Layer.SLICE, dict(loadTag=tag, sliceObj=slice_obj)))
        elif head == "X":
            m = _XNAME.search(tok)
            conf_list.append((Layer.EXPERIMENTAL, dict(name=m.group(1), conf=m.group(2), output_dim=output_dim)))
        elif head == "E":
            assert not embedding_defined
            conf_list[-1][-1]["isEmbedding"] = True
            embedding_defined = True
        elif head == "L":
            conf_list[-1][-1]["supervised"] = True
        elif head == "T":
            conf_list[-1][-1]["tag"] = body
        else:
            raise ValueError(f"Unknown layer config {tok} in network config {tokens}")
    return conf_list


def toNumpy(x):
    """reference models/__init__.py:153-161, for device tensors / SparseTensor."""
    from ..ops import SparseTensor
    if isinstance(x, SparseTensor):
        import numpy as np
        return {"indices": x.indices.cpu().numpy(), "values": x.values.cpu().numpy(),
                "dense_shape": np.asarray(x.dense_shape, dtype=np.int64)}
    return x.detach().cpu().numpy()


def add_subparsers(parser):
    """reference models/__init__.py:16-31: positional `model` argument + the model's own argument group."""
    import contextlib
    import importlib
    import os
    import pkgutil
    model_list = [name for _, name, _ in pkgutil.iter_modules(path=__path__) if not name.startswith("_")]
    parser.add_argument("model", choices=model_list, help="Network model selected for experiment")
    try:
        with open(os.devnull, "w") as null, contextlib.redirect_stderr(null):
            known, _ = parser.parse_known_args()
    except SystemExit:
        return
    model = importlib.import_module("." + known.model, package=__name__)
    if hasattr(model, "add_subparser_args"):
        model.add_subparser_args(parser)
        print(f"Using model: {model}")
