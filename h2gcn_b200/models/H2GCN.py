"""H2GCN forward pass on B200 — mirrors `h2gcn/models/H2GCN.py` of the reference (same constructor, call signature,
argv names, `args.objects` keys) with the aggregation path replaced by the fused CUDA kernels.

Two ways to execute the same layer list (reference: H2GCN.__init__ :210-292, H2GCN.call :294-346):

* interpreter — one layer object at a time, like the reference's Python loop; supports `returnBefore`,
  `executeAfter`, `addSupervision`, `saveActivations`.
* fused program — the layer list is compiled ONCE into a fixed launch sequence over a single zero-copy concat buffer
  (SURVEY.md §7.2): the first dense layer writes relu(X W0) into its final column slot, every `G-V` round is one fused
  launch that reads its input slot and writes all hops into their slots, `C`/`D`/`V`/`R`/`T` cost nothing, and the
  classifier reads the buffer in place.  Used by plain inference calls.
"""
import torch

from .. import ops
from . import Layer, parse_network_setup, toNumpy  # noqa: F401  (re-exported like the reference's `from . import *`)
from . import _layers as layers


def add_subparser_args(parser):
    """Same argv names and defaults as the reference (H2GCN.py:9-30)."""
    sub = parser.add_argument_group("H2GCN Model Arguments (H2GCN.py)")
    sub.add_argument("--network_setup", type=str, default="M64-R-T1-G-V-T2-G-V-C1-C2-D0.5-MO",
                     help="Default to H2GCN-2 (%(default)s)")
    sub.add_argument("--dropout", type=float, default=0.5, help="Default dropout rate")
    sub.add_argument("--hidden", type=int, default=64)
    sub.add_argument("--adj_nhood", default=["1", "2"], type=str, nargs="+")
    sub.add_argument("--optimizer", type=str, default="adam", help="(default: %(default)s)")
    sub.add_argument("--lr", type=float, default=0.01, help="(default: %(default)s)")
    sub.add_argument("--l2_regularize_weight", type=float, default=5e-4, help="(default: %(default)s)")
    sub.add_argument("--early_stopping", type=int, default=0,
                     help="Number of epochs used to decide early stopping (0 to disable) (default: %(default)s)")
    sub.add_argument("--best_val_criteria", choices=["val_acc", "val_loss"], default="val_acc")
    sub.add_argument("--save_activations", action="store_true")
    sub.add_argument("--save_predictions", nargs="+", type=bool, default=True)
    sub.add_argument("--no_feature_normalize", action="store_true")
    parser.function_hooks["argparse"].append(argparse_callback)


def argparse_callback(args):
    """H2GCN.py:33-43."""
    dataset = args.objects["dataset"]
    layer_setups = parse_network_setup(args.network_setup, dataset.num_labels, _dense_units=args.hidden,
                                       _dropout_rate=args.dropout, parse_preprocessing=True)
    if Layer.GCN in set(x[0] for x in layer_setups):
        preprocessing_data(args, getAdjNormHops=args.adj_nhood)
    else:
        preprocessing_data(args, getAdjHops=args.adj_nhood)
    initialize_model(args, layer_setups, args.optimizer, args.lr, args.l2_regularize_weight, args.early_stopping)


def preprocessing_data(args, **kwargs):
    """H2GCN.py:46-54: feature row-normalisation, removeEye, device tensors (adjacency powers on the GPU)."""
    dataset = args.objects["dataset"]
    if not args.no_feature_normalize:
        dataset.row_normalize_features()
    dataset.adj_remove_eye()
    args.objects["tensors"] = vars(dataset.getTensors(getDenseAdj=False, **kwargs))


def initialize_model(args, layer_setups, optimizer, lr, l2_regularize_weight, early_stopping):
    """H2GCN.py:57-206, inference part: registers the same `args.objects` keys.  The training loop (optimizer,
    checkpoints, early stopping) is outside the accelerated path."""
    model = H2GCN(layer_setups, l2_regularize_weight=l2_regularize_weight)

    state = dict(opt=None)

    def train_step(adj, adj_hops, features, y_train, train_mask, **kwargs):
        """H2GCN.py:66-74: forward (training=True), masked softmax-CE + l2, backward, optimizer step."""
        loss, grads = model.loss_and_grads(adj, features, adj_hops, y_train, train_mask)
        params = model.trainable_variables
        if state["opt"] is None:
            if str(optimizer).lower() != "adam":
                raise NotImplementedError("only --optimizer adam is wired up")
            state["opt"] = torch.optim.Adam(params, lr=lr, eps=1e-7)     # keras Adam defaults
        for w, g in zip(params, grads):
            w.grad = g
        state["opt"].step()
        return dict(train_loss=loss)

    def predict_step(adj, adj_hops, features, **kwargs):
        return model(adj, features, adj_hops, training=False)

    def test_step(adj, adj_hops, features, y_train, train_mask, y_val, val_mask, y_test, test_mask, **kwargs):
        pred = model(adj, features, adj_hops, training=False)

        def acc(y, m):  # masked accuracy, _metrics.py:18-25
            hit = (pred.argmax(1) == y.argmax(1)).float() * m
            return (hit.sum() / m.sum().clamp_min(1)).item()
        return dict(train_acc=acc(y_train, train_mask), val_acc=acc(y_val, val_mask), test_accuracy=acc(y_test, test_mask),
                    monitor=dict())

    def embed_step(adj, adj_hops, features, use_relu=False, **kwargs):
        return model.getEmbeddings(adj, features, adj_hops)

    args.objects["model"] = model
    args.objects["train_step"] = train_step
    args.objects["test_step"] = test_step
    args.objects["predict_step"] = predict_step
    args.objects["embed_step"] = embed_step


# ----------------------------------------------------------------------------------------------------------------------
class _FusedProgram:
    """Layer list -> fixed launch sequence over one concat buffer.  Built by symbolic execution of the list: every
    produced tensor is a list of (leaf id, width) parts; leaves are outputs of compute layers (dense[+relu], G[+V]).
    If the classifier input mentions each leaf at most once and every G input is a contiguous run of it, all leaves get
    their final column offsets up front and the concats vanish."""

    def __init__(self, model, adjhops, feat_dim, n_rows, device):
        self.ok = False
        self.dropout_feeds_classifier = []   # one flag per Dropout layer met by the symbolic execution
        objs = model.layer_objs
        leaves = {}      # leaf id -> dict(width, kind, ...)
        cur = None       # list of leaf ids (symbolic current tensor); None = the sparse input
        tagged = {}
        steps = []
        sparse = True
        for ind, layer in enumerate(objs):
            if isinstance(layer, layers.SparseDense) and sparse:
                lid = len(leaves)
                leaves[lid] = dict(width=layer.output_dim, kind="sparse_dense", layer=layer, relu=False)
                steps.append(lid)
                cur, sparse = [lid], False
            elif sparse:
                return  # SparseDropout / to_dense first: interpreter only
            elif isinstance(layer, layers.ReLU):
                if len(cur) != 1 or leaves[cur[0]].get("consumed") or leaves[cur[0]]["kind"] == "gcn" or \
                        any(cur[0] in t for t in tagged.values()):
                    return
                leaves[cur[0]]["relu"] = True  # fold into the producer's epilogue
            elif isinstance(layer, layers.GCNLayer):
                nh = len(layer.selected(adjhops))
                if nh == 0:
                    return
                win = sum(leaves[l]["width"] for l in cur)
                lid = len(leaves)
                leaves[lid] = dict(width=nh * win, kind="gcn", layer=layer, src=list(cur), d=win, nh=nh)
                for l in cur:
                    leaves[l]["consumed"] = True
                steps.append(lid)
                cur = [lid]
                if ind + 1 >= len(objs) or not isinstance(objs[ind + 1], layers.Flatten):
                    return  # a 3-D [N,H,d] tensor flows on: interpreter only
            elif isinstance(layer, layers.Flatten):
                pass
            elif isinstance(layer, layers.Dropout):
                # inference: identity.  Training applies ONE dropout, on the classifier input: remember where the
                # Dropout layers sit so that loss_and_grads can refuse any other placement instead of ignoring it.
                self.dropout_feeds_classifier.append(ind + 1 == len(objs) - 1 and isinstance(objs[-1], layers.Dense))
            elif isinstance(layer, layers.ConcatLayer):
                sel = [v for name, v in tagged.items() if name in layer.tags]
                cur = (list(cur) if layer.addInputs else []) + [l for v in sel for l in v]
            elif isinstance(layer, layers.Dense):
                if ind != len(objs) - 1:
                    return
                self.out_layer = layer
            else:
                return
            if ind in model.tagsDict and cur is not None:
                tagged[model.tagsDict[ind]] = list(cur)
        if sparse or not isinstance(objs[-1], layers.Dense) or len(set(cur)) != len(cur):
            return
        # column offsets: classifier input order first, leftovers after it
        order = list(cur) + [l for l in leaves if l not in cur]
        off, pos = {}, 0
        for l in order:
            pad = (-pos) % 4
            if pad and l in cur:
                return  # final layout would need padding inside the classifier input
            pos += pad
            off[l] = pos
            pos += leaves[l]["width"]
        for lid in steps:
            lf = leaves[lid]
            if lf["kind"] == "gcn":
                src = lf["src"]
                if any(off[b] != off[a] + leaves[a]["width"] for a, b in zip(src, src[1:])) or lf["d"] % 4 or off[src[0]] % 4:
                    return
        self.total_width = (pos + 3) // 4 * 4
        self.final_width = sum(leaves[l]["width"] for l in cur)
        self.leaves, self.steps, self.off = leaves, steps, off
        self.buf = torch.empty(n_rows, self.total_width, dtype=torch.float32, device=device)
        self._gbuf = None       # gradient twin of the concat buffer (training), allocated once
        self.adjhops = adjhops
        self.ok = True

    # ---- training: same launch sequence forward, mirrored offsets backward ------------------
    def train_forward(self, features, dropout_rate, generator=None):
        """Forward with dropout on the classifier input (keras Dropout, H2GCN.py:257); keeps what backward needs."""
        logits_in = self.run(features, classify=False)
        keep = 1.0 - dropout_rate
        if dropout_rate > 0:
            self._mask = (torch.rand(logits_in.shape, device=logits_in.device, generator=generator) < keep).float() / keep
            self._dropped = logits_in * self._mask
        else:
            self._mask, self._dropped = None, logits_in
        self._features = features
        return self.out_layer(self._dropped)

    def backward(self, dlogits, features_t):
        """Gradients of the kernels of the first dense layer and of the classifier for d(loss)/d(logits) = dlogits.
        The aggregation rounds run backwards through the SAME fused kernels: A_h is symmetric, so dX = sum_h A_h dY_h
        with hop h reading its own slice of the gradient buffer (`HopPlan.run_multi`) and `sum_slices` adding them."""
        W_out = self.out_layer.kernel
        # the two classifier-side contractions on the tcgen05 kernel (h2_dense_tc_f32, transposed forms): dW = final^T dlogits
        # and dfinal = dlogits W^T, the latter written straight into the gradient buffer
        gW_out = ops.matmul(self._dropped, dlogits, trans_a=True)            # [W, C]
        gbuf = self._gbuf
        if gbuf is None or gbuf.shape != self.buf.shape:
            gbuf = self._gbuf = torch.empty_like(self.buf)
        gbuf.zero_()
        ops.matmul(dlogits, W_out, trans_w=True, out=gbuf)                   # [N, W] into columns [0, final_width)
        if self._mask is not None:
            gbuf[:, :self.final_width].mul_(self._mask)
        gW0 = None
        for lid in reversed(self.steps):
            lf, o = self.leaves[lid], self.off[lid]
            if lf["kind"] == "gcn":
                plan = lf["layer"].plan_for(self.adjhops)
                so, d, nh = self.off[lf["src"][0]], lf["d"], lf["nh"]
                tmp = torch.empty(self.buf.shape[0], nh * d, dtype=torch.float32, device=self.buf.device)
                plan.run_multi(gbuf, [o + h * d for h in range(nh)], tmp, [h * d for h in range(nh)], d)
                ops.sum_slices(tmp, d, nh, gbuf[:, so:so + d], accumulate=True)
            else:
                p = lf["width"]
                g0 = gbuf[:, o:o + p]
                if lf["relu"]:
                    ops.sum_slices(g0, p, 0, g0, accumulate=True, mask_src=self.buf[:, o:o + p])   # g0 *= (r0 > 0)
                gW0 = ops.sparse_dense(features_t, g0.contiguous())        # X^T . dr0  ->  [F, p]
        return gW0, gW_out

    def capture(self, features):
        """The whole launch sequence of `run` as ONE CUDA graph: on Cora-sized graphs the forward is a
        handful of ~5-30 us kernels and the Python / launch overhead between them is a large part of the wall time.
        Returns the torch.cuda.CUDAGraph; `graph_out` holds the logits it writes.  The graph reads the SAME feature tensor,
        weights and concat buffer: replay after updating them in place."""
        for lid in self.steps:     # scratch must be bound before the capture (no allocation inside)
            lf = self.leaves[lid]
            if lf["kind"] == "gcn":
                lf["layer"].plan_for(self.adjhops).reserve(lf["d"])
        side = torch.cuda.Stream(device=self.buf.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self.run(features)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            self.graph_out = self.run(features)
        return graph

    def run(self, features, classify=True):
        buf = self.buf
        for lid in self.steps:
            lf, o = self.leaves[lid], self.off[lid]
            if lf["kind"] == "sparse_dense":
                lf["layer"](features, relu=lf["relu"], out=buf, out_col_off=o)
            else:
                plan = lf["layer"].plan_for(self.adjhops)
                so, d = self.off[lf["src"][0]], lf["d"]
                plan.run(buf[:, so:so + d], buf, [o + h * d for h in range(lf["nh"])], d=d)
        if not classify:
            return buf[:, :self.final_width]
        return self.out_layer(buf[:, :self.final_width])


class H2GCN:
    """reference H2GCN.py:209-367 (tf.keras.Model) — same attributes and call signature."""

    def __init__(self, layer_setups, sparse_input=True, l2_regularize_weight=0):
        self.layer_objs = []
        self.dropout_inds = []
        self.attention_inds = []
        self.supervised_inds = []
        self.graph_inds = []
        self.graph_hops_inds = []
        self.concat_inds = []
        self.experimental_inds = []
        self.embedding_ind = None
        self.output_ind = None
        self.tagsDict = dict()
        self.l2_regularize_weight = l2_regularize_weight
        self._fused = {}
        names = {}

        def add(obj):
            obj._assign_name(names)
            self.layer_objs.append(obj)

        for ind, (layerType, layerConf) in enumerate(layer_setups):
            layerConf = dict(layerConf)
            layerTag = layerConf.pop("tag", None)
            here = len(self.layer_objs)
            if layerType == Layer.DENSE:
                if layerConf.get("isEmbedding", False):
                    self.embedding_ind = here
                if layerConf.get("beginOutput", False):
                    self.output_ind = here
                if sparse_input:
                    add(layers.SparseDense(layerConf["units"], use_bias=layerConf["use_bias"],
                                           kernel_regularizer=l2_regularize_weight))
                    sparse_input = False
                else:
                    add(layers.Dense(layerConf["units"], use_bias=layerConf["use_bias"],
                                     kernel_regularizer=l2_regularize_weight))
            elif layerType == Layer.DROPOUT:
                self.dropout_inds.append(here)
                add(layers.SparseDropout(layerConf["dropout_rate"]) if sparse_input
                    else layers.Dropout(layerConf["dropout_rate"]))
            elif layerType == Layer.SLICE:
                self.concat_inds.append(here)
                add(layers.SliceLayer(**layerConf))
            elif layerType == Layer.IDENTITY:
                add(layers.ToDense())
                sparse_input = False
            elif layerType == Layer.GCN:
                self.graph_hops_inds.append(here)
                add(layers.GCNLayer(**layerConf))
            elif layerType == Layer.RELU:
                add(layers.ReLU())
            elif layerType == Layer.VECTORIZE:
                add(layers.Flatten())
            elif layerType == Layer.CONCAT:
                self.concat_inds.append(here)
                add(layers.ConcatLayer(tags=layerConf["tags"], addInputs=layerConf["addInputs"]))
            else:
                # The reference reaches `Layer.STOP_GRADIENT` here, which does not exist, and dies with AttributeError
                # (SURVEY.md §8a quirks); the intended ValueError is raised instead.
                raise ValueError(f"Unsupported layer type {layerType} specified in this model.")
            if layerConf.get("supervised", False):
                self.supervised_inds.append(len(self.layer_objs) - 1)
            if layerTag:
                if len(self.layer_objs) - 1 in self.tagsDict:
                    print(f"WARNING: overriding tag {layerTag} in layer {len(self.layer_objs) - 1}")
                self.tagsDict[len(self.layer_objs) - 1] = layerTag

    # ---- weights ---------------------------------------------------------------------------------------------------
    @property
    def trainable_variables(self):
        return [w for layer in self.layer_objs for w in getattr(layer, "weights", [])]

    def set_weights(self, arrays, device="cuda"):
        """Assign kernels/biases in layer order (what tf.train.Checkpoint restore does in the reference)."""
        it = iter(arrays)
        for layer in self.layer_objs:
            if isinstance(layer, (layers.SparseDense, layers.Dense)):
                layer.kernel = torch.as_tensor(next(it), dtype=torch.float32).to(device).contiguous()
                if layer.use_bias:
                    layer.bias = torch.as_tensor(next(it), dtype=torch.float32).to(device).contiguous()
        self._fused.clear()

    # ---- execution -------------------------------------------------------------------------------------------------
    def __call__(self, adj, inputs, adjhops, training=False, returnBefore=0, executeAfter=0, addSupervision=False,
                 saveActivations=None, **kwargs):
        return self.call(adj, inputs, adjhops, training, returnBefore, executeAfter, addSupervision, saveActivations,
                         **kwargs)

    def _fused_program(self, inputs, adjhops):
        key = (tuple(id(a) for a in adjhops), inputs.dense_shape, str(inputs.device))
        prog = self._fused.get(key)
        if prog is None:
            prog = self._fused[key] = _FusedProgram(self, adjhops, inputs.dense_shape[1], inputs.n_rows, inputs.device)
        return prog

    def call(self, adj, inputs, adjhops, training=False, returnBefore=0, executeAfter=0, addSupervision=False,
             saveActivations=None, **kwargs):
        plain = (not training and returnBefore == 0 and executeAfter == 0 and not addSupervision
                 and saveActivations is None and isinstance(inputs, ops.SparseTensor) and len(adjhops) > 0)
        if plain:
            prog = self._fused_program(inputs, adjhops)
            if prog.ok:
                return prog.run(inputs)
        return self._interpret(adj, inputs, adjhops, training, returnBefore, executeAfter, addSupervision,
                               saveActivations, **kwargs)

    def _interpret(self, adj, inputs, adjhops, training, returnBefore, executeAfter, addSupervision, saveActivations,
                   **kwargs):
        """The reference's interpreter loop (H2GCN.py:307-346), one C-ABI op per layer."""
        supervisedOutputs = []
        taggedOutputs = dict()
        if saveActivations is not None:
            saveActivations["inputs/inputs"] = toNumpy(inputs)
            saveActivations["inputs/adj"] = toNumpy(adj)
            for i in range(len(adjhops)):
                saveActivations[f"inputs/adjhops/{i}"] = toNumpy(adjhops[i])
        if returnBefore <= 0:
            returnBefore = len(self.layer_objs) + returnBefore
        if executeAfter < 0:
            executeAfter = len(self.layer_objs) + executeAfter
        for ind, layer in enumerate(self.layer_objs):
            if ind == returnBefore:
                return inputs
            elif ind < executeAfter:
                continue
            if ind in self.concat_inds:
                inputs = layer(inputs, **taggedOutputs)
            elif ind in self.graph_hops_inds:
                inputs = layer(adjhops, inputs)
            elif ind in self.dropout_inds:
                inputs = layer(inputs, training=training)
            else:
                inputs = layer(inputs)
            if addSupervision and (ind in self.supervised_inds):
                supervisedOutputs.append(self(adj, inputs, adjhops, training, executeAfter=self.output_ind,
                                              addSupervision=False, **kwargs))
            if saveActivations is not None:
                saveActivations[f"activations/{ind}-{layer.name}"] = toNumpy(inputs)
            if ind in self.tagsDict:
                taggedOutputs[self.tagsDict[ind]] = inputs
        if addSupervision:
            return inputs, supervisedOutputs
        return inputs

    def loss_and_grads(self, adj, inputs, adjhops, labels, mask, dropout_rate=None, generator=None):
        """Training forward + backward on the fused program.  Loss = masked softmax cross-entropy
        (h2gcn/models/_metrics.py:8-16) + sum_k l2 * ||W_k||^2 (keras.regularizers.l2, H2GCN.py:239-248, :363-367).
        Returns (loss, [dL/dW0, dL/dW_out]) in the order of `trainable_variables`."""
        prog = self._fused_program(inputs, adjhops)
        if not prog.ok:
            raise NotImplementedError("training is implemented for the fused layer-list family only")
        if any(getattr(l, "use_bias", False) for l in self.layer_objs):
            raise NotImplementedError("training with bias terms (F layers) is not wired up")
        if not all(prog.dropout_feeds_classifier):
            # e.g. M64-R-D0.5-T1-G-V-...: keras would drop activations in the middle of the network; the fused
            # training path only implements the H2GCN placement (Dropout directly in front of the final Dense)
            raise NotImplementedError("training: a Dropout layer that does not directly feed the final Dense layer is "
                                      "not implemented on the fused path")
        if dropout_rate is None:
            rates = [l.rate for l in self.layer_objs if isinstance(l, layers.Dropout)]
            dropout_rate = rates[-1] if rates else 0.0
        key = id(inputs)
        if getattr(self, "_feat_t_key", None) != key:      # X^T as CSR, once per feature tensor (for dW0 = X^T dr0)
            xt = inputs.to_scipy().T.tocsr()
            self._feat_t = ops.SparseTensor.from_scipy(xt, inputs.device)
            self._feat_t_key = key
        logits = prog.train_forward(inputs, dropout_rate, generator)
        m = mask / mask.sum()
        logp = torch.log_softmax(logits, dim=1)
        ce = -(labels * logp).sum(1)
        l2 = self.l2_regularize_weight
        params = self.trainable_variables
        loss = (ce * m).sum() + sum(l2 * (w * w).sum() for w in params)
        dlogits = (torch.softmax(logits, dim=1) * labels.sum(1, keepdim=True) - labels) * m[:, None]
        gW0, gW_out = prog.backward(dlogits.contiguous(), self._feat_t)
        grads = [gW0 + 2 * l2 * params[0], gW_out + 2 * l2 * params[1]]
        return float(loss), grads

    def callOutputNetwork(self, adj, inputs, adjhops, training=False, returnBefore=0, **kwargs):
        return self(adj, inputs, adjhops, training, returnBefore, self.output_ind, **kwargs)

    def getEmbeddings(self, adj, inputs, adjhops, **kwargs):
        return self(adj, inputs, adjhops, training=False, returnBefore=self.embedding_ind + 1)
