"""Layer objects of the B200-native H2GCN path — mirrors `h2gcn/models/_layers.py` of the reference (same class
names, constructor arguments and call signatures); tensors are torch CUDA tensors / ops.SparseTensor and every
arithmetic op is a kernel of libh2gcn_b200.so (no TensorFlow, no CPU fallback).

These classes are the layer-by-layer ("interpreter") form used when a caller asks for intermediate activations or
partial execution.  The fused zero-copy form of the same layer lists lives in H2GCN.py (`_FusedProgram`).
"""
import math

import torch

from .. import ops


class _Named:
    """Keras-style layer names (`gcn_layer`, `gcn_layer_1`, ...) so that saveActivations keys match the reference's
    (`activations/{ind}-{layer.name}`, H2GCN.py:333-337).  The counter is per model, see H2GCN.__init__."""
    base_name = "layer"

    def _assign_name(self, counters):
        k = counters.get(self.base_name, 0)
        counters[self.base_name] = k + 1
        self.name = self.base_name if k == 0 else f"{self.base_name}_{k}"


def glorot_uniform(fan_in, fan_out, device, generator=None):
    """Keras default kernel initialiser (add_weight without initializer, _layers.py:31-35)."""
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    w = torch.rand(fan_in, fan_out, device=device, dtype=torch.float32, generator=generator)
    return w.mul_(2 * limit).sub_(limit)


class SparseDropout(_Named):
    """_layers.py:7-19.  Inference is the identity; training keeps each stored value with prob 1-p and rescales."""
    base_name = "sparse_dropout"

    def __init__(self, drop_prob):
        self.drop_prob = drop_prob

    def __call__(self, input, training=False):
        if not training or self.drop_prob == 0:
            return input
        keep = 1 - self.drop_prob
        mask = torch.floor(keep + torch.rand_like(input.values)).bool()
        # tf.sparse.retain drops the masked entries; zeroing them gives the same products
        return ops.SparseTensor(input.rowptr, input.col, torch.where(mask, input.values / keep, 0.0), input.dense_shape)


class SparseDense(_Named):
    """_layers.py:22-52: kernel [F, output_dim] (+bias) applied to a sparse input."""
    base_name = "sparse_dense"

    def __init__(self, output_dim, use_bias=False, activation=None, kernel_regularizer=None):
        self.output_dim = output_dim
        self.activation = activation
        self.use_bias = use_bias
        self.kernel_regularizer = kernel_regularizer
        self.kernel = None
        self.bias = None

    def build(self, input_shape, device="cuda", generator=None):
        self.kernel = glorot_uniform(int(input_shape[-1]), self.output_dim, device, generator)
        if self.use_bias:
            self.bias = torch.zeros(self.output_dim, device=device, dtype=torch.float32)

    @property
    def weights(self):
        return [w for w in (self.kernel, self.bias) if w is not None]

    # a feature matrix at least this dense is a true dense contraction (syn-products / ogbn features):
    # it is densified once and goes through the tensor-core kernel instead of the CSR gather
    DENSE_THRESHOLD = 0.25

    def __call__(self, input, relu=False, out=None, out_col_off=0):
        if self.kernel is None:
            self.build(input.shape, input.device)
        n, f = input.dense_shape
        if n and f and input.nnz >= self.DENSE_THRESHOLD * n * f:
            xd = input.__dict__.get("_dense")
            if xd is None:
                xd = input.__dict__["_dense"] = ToDense()(input)
            y = ops.matmul(xd, self.kernel, bias=self.bias, relu=relu, out=out, out_col_off=out_col_off)
        else:
            y = ops.sparse_dense(input, self.kernel, self.bias, relu=relu, out=out, out_col_off=out_col_off)
        if self.activation:
            y = self.activation(y)
        return y


class Dense(_Named):
    """keras.layers.Dense as used at H2GCN.py:244-249."""
    base_name = "dense"

    def __init__(self, units, use_bias=True, kernel_regularizer=None):
        self.units = units
        self.use_bias = use_bias
        self.kernel_regularizer = kernel_regularizer
        self.kernel = None
        self.bias = None

    def build(self, input_shape, device="cuda", generator=None):
        self.kernel = glorot_uniform(int(input_shape[-1]), self.units, device, generator)
        if self.use_bias:
            self.bias = torch.zeros(self.units, device=device, dtype=torch.float32)

    @property
    def weights(self):
        return [w for w in (self.kernel, self.bias) if w is not None]

    def __call__(self, inputs, relu=False):
        if self.kernel is None:
            self.build(inputs.shape, inputs.device)
        return ops.dense(inputs, self.kernel, self.bias, relu=relu)


class ReLU(_Named):
    base_name = "re_lu"

    def __call__(self, inputs):
        return ops.relu_slice(inputs, relu=True)


class Flatten(_Named):
    """keras Flatten for the `V` token (H2GCN.py:271-272): [N, H, d] -> [N, H*d], a view."""
    base_name = "flatten"

    def __call__(self, inputs):
        return inputs.reshape(inputs.shape[0], -1)


class Dropout(_Named):
    """keras Dropout (H2GCN.py:257).  Identity at inference — the only mode on the accelerated path."""
    base_name = "dropout"

    def __init__(self, rate):
        self.rate = rate

    def __call__(self, inputs, training=False):
        if not training or self.rate == 0:
            return inputs
        return torch.nn.functional.dropout(inputs, self.rate, training=True)  # training loop is out of scope


class GCNLayer(_Named):
    """_layers.py:54-81.  call(adjhops, inputs) -> [N, H, d]: hop h of the (sub)set lands in [:, h, :].

    All selected hops are computed by ONE fused launch (ops.HopPlan); the reference's per-hop
    tf.sparse.sparse_dense_matmul calls, tf.stack and the GPU column-split workaround (:65-74, a TF 2^31-element
    limit) have no counterpart here — the kernel uses 64-bit offsets."""
    SIGNATURE = ["adjhops", "inputs"]
    base_name = "gcn_layer"

    def __init__(self, hops=None):
        self.hops = hops
        self.cpu_large_spmatmul = False

    def selected(self, adjhops):
        return [x for ind, x in enumerate(adjhops) if (self.hops is None or ind in self.hops)]

    # The plan (schedule, tile-bitmap format, scratch) is built once per graph and shared by every GCNLayer that
    # aggregates over the same hop tensors (both rounds of H2GCN-2).  It is stored ON the first hop tensor, so it lives
    # exactly as long as the graph's tensors do (no process-global cache pinning device memory), and the handle
    # serialises rounds issued from different streams (include/h2gcn_b200.h, h2_graph_bind_workspace).
    def plan_for(self, adjhops):
        sel = self.selected(adjhops)
        cache = sel[0].__dict__.setdefault("_hop_plans", {})
        key = tuple(id(x) for x in sel)
        entry = cache.get(key)
        if entry is None:
            entry = cache[key] = (sel[1:], ops.HopPlan(sel))     # keeps the other hop tensors alive: ids stay valid
        return entry[1]

    def sparse_dense_matmul(self, sp_a, b, ind=""):
        """Single-hop product, kept for API parity (_layers.py:62-76)."""
        out = torch.empty(sp_a.n_rows, b.shape[1], dtype=torch.float32, device=b.device)
        return ops.HopPlan([sp_a]).run(b, out, [0])

    def __call__(self, adjhops, inputs):
        plan = self.plan_for(adjhops)
        n, d = plan.n_rows, inputs.shape[1]
        pad = (-d) % 4
        x = inputs
        if pad or x.stride(0) % 4 or x.data_ptr() % 16:
            x = torch.zeros(inputs.shape[0], d + pad, dtype=torch.float32, device=inputs.device)
            ops.relu_slice(inputs, out=x, relu=False)
        dd = d + pad
        H = len(plan.hops)
        out = torch.empty(n, H, dd, dtype=torch.float32, device=inputs.device)
        plan.run(x, out.view(n, H * dd), [h * dd for h in range(H)], d=dd)
        return out[:, :, :d] if pad else out


class ConcatLayer(_Named):
    """_layers.py:83-96: concat([inputs] + [tagged[t] for t in kwargs order if t in tags], axis)."""
    base_name = "concat_layer"

    def __init__(self, tags, axis=-1, addInputs=True):
        self.tags = tags
        self.axis = axis
        self.addInputs = addInputs

    def __call__(self, *args, **kwargs):
        selected = [value for name, value in kwargs.items() if name in self.tags]
        parts = (list(args) + selected) if self.addInputs else selected
        if self.axis not in (-1, 1) or any(p.dim() != 2 for p in parts):
            return torch.cat(parts, self.axis)
        n = parts[0].shape[0]
        out = torch.empty(n, sum(p.shape[1] for p in parts), dtype=torch.float32, device=parts[0].device)
        off = 0
        for p in parts:
            ops.relu_slice(p, out=out[:, off:off + p.shape[1]], relu=False)
            off += p.shape[1]
        return out


class SumLayer(_Named):
    base_name = "sum_layer"

    def __init__(self, dim=-2):
        self.dim = dim

    def __call__(self, inputs):
        return inputs.sum(self.dim)


class SliceLayer(_Named):
    """_layers.py:108-117."""
    base_name = "slice_layer"

    def __init__(self, loadTag, sliceObj, **kwargs):
        self.tag = loadTag
        self.sliceObj = sliceObj

    def __call__(self, inputs, **kwargs):
        if self.tag:
            inputs = kwargs[self.tag]
        return inputs[:, self.sliceObj]


class ToDense(_Named):
    """tf.sparse.to_dense for the `I` token (H2GCN.py:262-264)."""
    base_name = "to_dense"
    __name__ = "to_dense"

    def __call__(self, sp):
        out = torch.zeros(sp.n_rows, sp.dense_shape[1], dtype=torch.float32, device=sp.device)
        idx = sp.indices
        out[idx[:, 0] - sp.row_begin, idx[:, 1]] = sp.values
        return out


experimentalDict = {}
