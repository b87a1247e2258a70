"""ENet parameter inventory, seeded synthetic weights and the flat weight container.

The reference never states the network: ``models.py:21-31`` imports a frozen
GraphDef (``pretrained_models/enet.pb``, derived from ``pretrained_models/model.h5``
through ``utils.py:49-83``) and both blobs are absent from the snapshot
(``.MISSING_LARGE_BLOBS``).  The tensor names ``input0`` / ``CATkrIDy``
(``models.py:15-16``) are pytorch2keras short names, i.e. the model was a PyTorch
ENet; this module therefore uses the canonical ENet of Paszke et al. 2016 with the
parameter naming of the widely used PyTorch implementation, so that a real
``state_dict`` can be packed with ``pack_flat`` unchanged.

Flat container ("BCENETW1"), little endian:
    char[8]  magic = b"BCENETW1"
    u32      n_tensors
    u32      num_classes
    f32      bn_eps
    u32      reserved (0)
    n_tensors x { char[96] name (NUL padded); u32 ndim; u32 dims[4]; u64 offset }
    raw fp32 data (offsets are relative to the start of the data section,
    which begins at the first 64-byte boundary after the table)
A missing ``*.weight`` for an activation means ReLU (decoder_relu=True in the
canonical model); a 1-element activation weight is a shared PReLU slope, a
C-element one is per-channel.  A ``<conv>.bias`` tensor next to a ``<conv>.weight`` is a
convolution bias (folded into the batch-norm shift at load time).

The network GRAPH travels in the container too (real blobs of this family differ in details,
SURVEY.md section 7 hard part 1), as two reserved tensors:
    ``__graph__``  float32 [n_blocks][7]: kind (0 down, 1 regular/dilated, 2 asymmetric, 3 up), stage,
                   index, cin, cout, internal width, dilation -- the block's parameter prefix is
                   ``{downsample|regular|dilated|asymmetric|upsample}{stage}_{index}``
    ``__spec__``   float32 [4]: initial max-pool kernel (3: 3x3 s2 p1 as PyTorch-ENet, 2: 2x2 s2 as the
                   paper / Keras ports), head kernel (3: ConvTranspose 3x3 s2 p1 op1, 2: 2x2 s2), 0, 0
Absent tensors mean the canonical graph (ENET_BLOCKS) and (3, 3).  The C loader validates the list
against what its kernels implement and every parameter shape against the list.
"""
import struct

import numpy as np

MAGIC = b"BCENETW1"
NAME_LEN = 96
BN_EPS = 1e-5

# (name, kind, args): the canonical encoder/decoder sequence
#   kind: "down" (cin, cout) | "reg" (ch, dilation) | "asym" (ch) | "up" (cin, cout)
ENET_BLOCKS = [
    ("downsample1_0", "down", (16, 64)),
    ("regular1_1", "reg", (64, 1)), ("regular1_2", "reg", (64, 1)),
    ("regular1_3", "reg", (64, 1)), ("regular1_4", "reg", (64, 1)),
    ("downsample2_0", "down", (64, 128)),
    ("regular2_1", "reg", (128, 1)), ("dilated2_2", "reg", (128, 2)),
    ("asymmetric2_3", "asym", (128,)), ("dilated2_4", "reg", (128, 4)),
    ("regular2_5", "reg", (128, 1)), ("dilated2_6", "reg", (128, 8)),
    ("asymmetric2_7", "asym", (128,)), ("dilated2_8", "reg", (128, 16)),
    ("regular3_0", "reg", (128, 1)), ("dilated3_1", "reg", (128, 2)),
    ("asymmetric3_2", "asym", (128,)), ("dilated3_3", "reg", (128, 4)),
    ("regular3_4", "reg", (128, 1)), ("dilated3_5", "reg", (128, 8)),
    ("asymmetric3_6", "asym", (128,)), ("dilated3_7", "reg", (128, 16)),
    ("upsample4_0", "up", (128, 64)),
    ("regular4_1", "reg", (64, 1)), ("regular4_2", "reg", (64, 1)),
    ("upsample5_0", "up", (64, 16)),
    ("regular5_1", "reg", (16, 1)),
]
DECODER_BLOCKS = {"upsample4_0", "regular4_1", "regular4_2", "upsample5_0", "regular5_1"}


KIND_CODE = {"down": 0, "reg": 1, "asym": 2, "up": 3}


def graph_rows(blocks=None, down_internal="in/4"):
    """ENET_BLOCKS-style list -> float32 [n][7] rows of the container's ``__graph__`` tensor"""
    rows = []
    for name, kind, args in (ENET_BLOCKS if blocks is None else blocks):
        import re
        m = re.match(r"[a-z]+(\d+)_(\d+)$", name)
        stage, index = int(m.group(1)), int(m.group(2))
        if kind in ("down", "up"):
            cin, cout = args
            ci = (cout if down_internal == "out/4" and kind == "down" else cin) // 4
            dil = 1
        else:
            cin = cout = args[0]
            ci = cin // 4
            dil = args[1] if kind == "reg" else 1
        rows.append([KIND_CODE[kind], stage, index, cin, cout, ci, dil])
    return np.asarray(rows, np.float32)


def block_name(kind_code, stage, index, dilation):
    base = {0: "downsample", 1: "regular" if dilation == 1 else "dilated", 2: "asymmetric", 3: "upsample"}[int(kind_code)]
    return f"{base}{int(stage)}_{int(index)}"


def _bn(prefix, c):
    return [(prefix + ".weight", (c,), "bn_gamma"), (prefix + ".bias", (c,), "bn_beta"),
            (prefix + ".running_mean", (c,), "bn_mean"), (prefix + ".running_var", (c,), "bn_var")]


def enet_param_spec(num_classes=15, encoder_relu=False, decoder_relu=True, prelu_per_channel=False, conv_bias=False,
                    head_kernel=3, blocks=None):
    """Ordered list of (name, shape, kind) for every parameter of the network."""
    spec = []
    last_ch = [16]

    def act(name, relu, ch=None):
        if not relu:
            spec.append((name + ".weight", ((ch if ch else last_ch[0]) if prelu_per_channel else 1,), "prelu"))

    def conv(name, shape, kind="conv"):
        spec.append((name + ".weight", shape, kind))
        cout = shape[1] if kind == "tconv" else shape[0]
        if conv_bias and name != "transposed_conv":
            spec.append((name + ".bias", (cout,), "bias"))
        last_ch[0] = cout

    conv("initial_block.main_branch", (13, 3, 3, 3))
    spec.extend(_bn("initial_block.batch_norm", 16))
    act("initial_block.out_activation", encoder_relu, 16)
    for name, kind, args in (ENET_BLOCKS if blocks is None else blocks):
        relu = decoder_relu if name in DECODER_BLOCKS else encoder_relu
        if kind == "down":
            cin, cout = args
            ci = cin // 4
            conv(f"{name}.ext_conv1.0", (ci, cin, 2, 2))
            spec.extend(_bn(f"{name}.ext_conv1.1", ci)); act(f"{name}.ext_conv1.2", relu)
            conv(f"{name}.ext_conv2.0", (ci, ci, 3, 3))
            spec.extend(_bn(f"{name}.ext_conv2.1", ci)); act(f"{name}.ext_conv2.2", relu)
            conv(f"{name}.ext_conv3.0", (cout, ci, 1, 1))
            spec.extend(_bn(f"{name}.ext_conv3.1", cout)); act(f"{name}.ext_conv3.2", relu)
            act(f"{name}.out_activation", relu, cout)
        elif kind in ("reg", "asym"):
            ch = args[0]
            ci = ch // 4
            conv(f"{name}.ext_conv1.0", (ci, ch, 1, 1))
            spec.extend(_bn(f"{name}.ext_conv1.1", ci)); act(f"{name}.ext_conv1.2", relu)
            if kind == "reg":
                conv(f"{name}.ext_conv2.0", (ci, ci, 3, 3))
                spec.extend(_bn(f"{name}.ext_conv2.1", ci)); act(f"{name}.ext_conv2.2", relu)
            else:
                conv(f"{name}.ext_conv2.0", (ci, ci, 5, 1))
                spec.extend(_bn(f"{name}.ext_conv2.1", ci)); act(f"{name}.ext_conv2.2", relu)
                conv(f"{name}.ext_conv2.3", (ci, ci, 1, 5))
                spec.extend(_bn(f"{name}.ext_conv2.4", ci)); act(f"{name}.ext_conv2.5", relu)
            conv(f"{name}.ext_conv3.0", (ch, ci, 1, 1))
            spec.extend(_bn(f"{name}.ext_conv3.1", ch)); act(f"{name}.ext_conv3.2", relu)
            act(f"{name}.out_activation", relu, ch)
        elif kind == "up":
            cin, cout = args
            ci = cin // 4
            conv(f"{name}.main_conv1.0", (cout, cin, 1, 1))
            spec.extend(_bn(f"{name}.main_conv1.1", cout))
            conv(f"{name}.ext_conv1.0", (ci, cin, 1, 1))
            spec.extend(_bn(f"{name}.ext_conv1.1", ci)); act(f"{name}.ext_conv1.2", relu)
            conv(f"{name}.ext_tconv1", (ci, ci, 2, 2), "tconv")
            spec.extend(_bn(f"{name}.ext_tconv1_bnorm", ci)); act(f"{name}.ext_tconv1_activation", relu)
            conv(f"{name}.ext_conv2.0", (cout, ci, 1, 1))
            spec.extend(_bn(f"{name}.ext_conv2.1", cout))
            act(f"{name}.out_activation", relu, cout)
    spec.append(("transposed_conv.weight", (16, num_classes, head_kernel, head_kernel), "tconv"))
    return spec


def synthetic_weights(seed=42, num_classes=15, encoder_relu=False, decoder_relu=True, **variant):
    """Seeded generator: He-scaled conv weights,
    BN gamma~U(0.5,1.5), beta~N(0,0.1), mean~N(0,0.1), var~U(0.5,1.5), PReLU 0.25
    (per-channel variant: U(0.05, 0.45)), conv biases N(0, 0.1).  ``variant``: see enet_param_spec."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape, kind in enet_param_spec(num_classes, encoder_relu, decoder_relu, **variant):
        if kind == "conv":
            fan_in = shape[1] * shape[2] * shape[3]
            w = rng.standard_normal(shape) * np.sqrt(2.0 / fan_in)
        elif kind == "tconv":   # (in, out, kh, kw); each output sees ~in*kh*kw/stride^2 taps
            fan_in = shape[0] * shape[2] * shape[3] / 4.0
            w = rng.standard_normal(shape) * np.sqrt(2.0 / fan_in)
        elif kind == "bn_gamma":
            w = rng.uniform(0.5, 1.5, shape)
        elif kind == "bn_beta" or kind == "bn_mean":
            w = rng.standard_normal(shape) * 0.1
        elif kind == "bn_var":
            w = rng.uniform(0.5, 1.5, shape)
        elif kind == "prelu":
            w = np.full(shape, 0.25) if shape == (1,) else rng.uniform(0.05, 0.45, shape)
        elif kind == "bias":
            w = rng.standard_normal(shape) * 0.1
        else:
            raise ValueError(kind)
        out[name] = np.ascontiguousarray(w, dtype=np.float32)
    return out


def pack_flat(weights, num_classes=None, bn_eps=BN_EPS, graph=None, initial_pool=None, head_kernel=None):
    """dict name -> float32 ndarray (<= 4 dims)  ->  bytes of the flat container.  ``graph`` (graph_rows),
    ``initial_pool`` and ``head_kernel`` add the reserved ``__graph__`` / ``__spec__`` tensors."""
    weights = dict(weights)
    if graph is not None:
        weights["__graph__"] = np.asarray(graph, np.float32)
    if initial_pool is not None or head_kernel is not None:
        weights["__spec__"] = np.asarray([initial_pool or 3, head_kernel or 3, 0, 0], np.float32)
    names = list(weights.keys())
    if num_classes is None:
        num_classes = int(weights["transposed_conv.weight"].shape[1])
    table = bytearray()
    chunks = []
    off = 0
    for n in names:
        a = np.ascontiguousarray(np.asarray(weights[n]), dtype="<f4")
        if a.ndim > 4:
            raise ValueError(f"{n}: more than 4 dims")
        nb = n.encode()
        if len(nb) >= NAME_LEN:
            raise ValueError(f"{n}: name too long")
        dims = list(a.shape) + [0] * (4 - a.ndim)
        table += struct.pack(f"<{NAME_LEN}sI4IQ", nb, a.ndim, *dims, off)
        chunks.append(a.tobytes())
        off += a.nbytes
    head = MAGIC + struct.pack("<IIfI", len(names), num_classes, float(bn_eps), 0)
    blob = bytearray(head + bytes(table))
    blob += b"\0" * ((-len(blob)) % 64)
    for c in chunks:
        blob += c
    return bytes(blob)


def unpack_flat(blob):
    """bytes -> (dict name -> float32 ndarray, num_classes, bn_eps)."""
    if blob[:8] != MAGIC:
        raise ValueError("not a BCENETW1 container")
    n, num_classes, eps, _ = struct.unpack_from("<IIfI", blob, 8)
    pos = 24
    ent = struct.calcsize(f"<{NAME_LEN}sI4IQ")
    data0 = pos + n * ent
    data0 += (-data0) % 64
    out = {}
    for _ in range(n):
        nb, ndim, d0, d1, d2, d3, off = struct.unpack_from(f"<{NAME_LEN}sI4IQ", blob, pos)
        pos += ent
        shape = (d0, d1, d2, d3)[:ndim]
        cnt = int(np.prod(shape)) if ndim else 1
        a = np.frombuffer(blob, dtype="<f4", count=cnt, offset=data0 + off).reshape(shape)
        out[nb.rstrip(b"\0").decode()] = a.copy()
    return out, num_classes, eps
