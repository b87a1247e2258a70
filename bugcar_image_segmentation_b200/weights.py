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
C-element one is per-channel.
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


def _bn(prefix, c):
    return [(prefix + ".weight", (c,), "bn_gamma"), (prefix + ".bias", (c,), "bn_beta"),
            (prefix + ".running_mean", (c,), "bn_mean"), (prefix + ".running_var", (c,), "bn_var")]


def enet_param_spec(num_classes=15, encoder_relu=False, decoder_relu=True):
    """Ordered list of (name, shape, kind) for every parameter of the network."""
    spec = []

    def act(name, relu):
        if not relu:
            spec.append((name + ".weight", (1,), "prelu"))

    spec.append(("initial_block.main_branch.weight", (13, 3, 3, 3), "conv"))
    spec.extend(_bn("initial_block.batch_norm", 16))
    act("initial_block.out_activation", encoder_relu)
    for name, kind, args in ENET_BLOCKS:
        relu = decoder_relu if name in DECODER_BLOCKS else encoder_relu
        if kind == "down":
            cin, cout = args
            ci = cin // 4
            spec.append((f"{name}.ext_conv1.0.weight", (ci, cin, 2, 2), "conv"))
            spec.extend(_bn(f"{name}.ext_conv1.1", ci)); act(f"{name}.ext_conv1.2", relu)
            spec.append((f"{name}.ext_conv2.0.weight", (ci, ci, 3, 3), "conv"))
            spec.extend(_bn(f"{name}.ext_conv2.1", ci)); act(f"{name}.ext_conv2.2", relu)
            spec.append((f"{name}.ext_conv3.0.weight", (cout, ci, 1, 1), "conv"))
            spec.extend(_bn(f"{name}.ext_conv3.1", cout)); act(f"{name}.ext_conv3.2", relu)
            act(f"{name}.out_activation", relu)
        elif kind in ("reg", "asym"):
            ch = args[0]
            ci = ch // 4
            spec.append((f"{name}.ext_conv1.0.weight", (ci, ch, 1, 1), "conv"))
            spec.extend(_bn(f"{name}.ext_conv1.1", ci)); act(f"{name}.ext_conv1.2", relu)
            if kind == "reg":
                spec.append((f"{name}.ext_conv2.0.weight", (ci, ci, 3, 3), "conv"))
                spec.extend(_bn(f"{name}.ext_conv2.1", ci)); act(f"{name}.ext_conv2.2", relu)
            else:
                spec.append((f"{name}.ext_conv2.0.weight", (ci, ci, 5, 1), "conv"))
                spec.extend(_bn(f"{name}.ext_conv2.1", ci)); act(f"{name}.ext_conv2.2", relu)
                spec.append((f"{name}.ext_conv2.3.weight", (ci, ci, 1, 5), "conv"))
                spec.extend(_bn(f"{name}.ext_conv2.4", ci)); act(f"{name}.ext_conv2.5", relu)
            spec.append((f"{name}.ext_conv3.0.weight", (ch, ci, 1, 1), "conv"))
            spec.extend(_bn(f"{name}.ext_conv3.1", ch)); act(f"{name}.ext_conv3.2", relu)
            act(f"{name}.out_activation", relu)
        elif kind == "up":
            cin, cout = args
            ci = cin // 4
            spec.append((f"{name}.main_conv1.0.weight", (cout, cin, 1, 1), "conv"))
            spec.extend(_bn(f"{name}.main_conv1.1", cout))
            spec.append((f"{name}.ext_conv1.0.weight", (ci, cin, 1, 1), "conv"))
            spec.extend(_bn(f"{name}.ext_conv1.1", ci)); act(f"{name}.ext_conv1.2", relu)
            spec.append((f"{name}.ext_tconv1.weight", (ci, ci, 2, 2), "tconv"))
            spec.extend(_bn(f"{name}.ext_tconv1_bnorm", ci)); act(f"{name}.ext_tconv1_activation", relu)
            spec.append((f"{name}.ext_conv2.0.weight", (cout, ci, 1, 1), "conv"))
            spec.extend(_bn(f"{name}.ext_conv2.1", cout))
            act(f"{name}.out_activation", relu)
    spec.append(("transposed_conv.weight", (16, num_classes, 3, 3), "tconv"))
    return spec


def synthetic_weights(seed=42, num_classes=15, encoder_relu=False, decoder_relu=True):
    """Seeded generator: He-scaled conv weights,
    BN gamma~U(0.5,1.5), beta~N(0,0.1), mean~N(0,0.1), var~U(0.5,1.5), PReLU 0.25."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape, kind in enet_param_spec(num_classes, encoder_relu, decoder_relu):
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
            w = np.full(shape, 0.25)
        else:
            raise ValueError(kind)
        out[name] = np.ascontiguousarray(w, dtype=np.float32)
    return out


def pack_flat(weights, num_classes=None, bn_eps=BN_EPS):
    """dict name -> float32 ndarray (<= 4 dims)  ->  bytes of the flat container."""
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
