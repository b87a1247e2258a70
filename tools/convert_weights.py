"""Pack real ENet weights into the BCENETW1 container that ``bc_load_enet`` reads.

    python tools/convert_weights.py model.pth  enet.bcw [--classes 15] [--bn-eps 1e-5]
    python tools/convert_weights.py model.h5   enet.bcw

The reference loads ``pretrained_models/enet.pb`` (models.py:24), frozen by ``utils.freeze_session``
(utils.py:49-83) from ``pretrained_models/model.h5``, which pytorch2keras produced from a PyTorch ENet
(tensor names ``input0`` / ``CAT*``, models.py:15-16).  None of the three files is in the reference's
source tree, so:

  * ``.pth`` / ``.pt`` (a PyTorch ``state_dict`` of the canonical ENet, or a checkpoint holding one
    under ``state_dict``): every tensor is looked up by its parameter name and shape-checked against
    ``weights.enet_param_spec``; a missing activation weight means ReLU; PReLU slopes may be shared or
    per channel; ``<conv>.bias`` tensors (checkpoints built with bias=True) are carried over and folded
    into the batch-norm shift by the loader, except on the class head (``transposed_conv.bias``: not
    implemented, refused); the head kernel (3x3 or 2x2) is taken from the tensor's shape; any other
    unknown key is reported.  Exercised by tests/test_abi_host.py.
  * ``.h5`` (Keras): pytorch2keras renames every layer to a random short name, so tensors cannot be
    matched by name.  They are matched by ORDER within each kind (conv / transposed-conv kernels, BN
    quadruples, PReLU slopes in graph order) and by shape, with kernels transposed from Keras'
    (kh, kw, in, out) to PyTorch's (out, in, kh, kw).  This needs ``h5py`` and a real ``model.h5``;
    neither exists in the build container, so this branch has never been run -- it fails loudly
    on the first shape that does not fit instead of guessing.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bugcar_image_segmentation_b200 import weights as W   # noqa: E402


def from_state_dict(sd, num_classes, encoder_relu=False, decoder_relu=True):
    if "state_dict" in sd and not any(k.endswith(".weight") for k in sd):
        sd = sd["state_dict"]
    sd = {k[7:] if k.startswith("module.") else k: v for k, v in sd.items()}
    tc = sd.get("transposed_conv.weight")
    head_kernel = int(tc.shape[-1]) if tc is not None and int(tc.shape[-1]) in (2, 3) else 3
    spec = W.enet_param_spec(num_classes, encoder_relu, decoder_relu, conv_bias=True, head_kernel=head_kernel)
    known = {n for n, _, _ in spec} | {"transposed_conv.bias"}
    extra = sorted(k for k in sd if k not in known and not k.endswith("num_batches_tracked"))
    if extra:
        print(f"warning: {len(extra)} tensors of the state_dict are not part of the ENet graph and are ignored: "
              + ", ".join(extra[:8]) + (" ..." if len(extra) > 8 else ""), file=sys.stderr)
    if "transposed_conv.bias" in sd and np.any(np.asarray(sd["transposed_conv.bias"]) != 0):
        raise ValueError("transposed_conv.bias is non-zero: a bias on the class head is not implemented by the loader")
    out = {}
    for name, shape, kind in spec:
        if kind == "bias" and name not in sd:
            continue                          # bias=False convolution (the canonical model)
        if name not in sd:
            if kind == "prelu":
                continue                      # ReLU variant of this activation: no slope tensor
            raise KeyError(f"state_dict has no tensor {name!r}")
        a = np.asarray(sd[name].detach().cpu().numpy() if hasattr(sd[name], "detach") else sd[name], dtype=np.float32)
        if kind == "prelu":
            if a.size not in (1, None) and a.ndim != 1:
                raise ValueError(f"{name}: PReLU slope must be 1-D")
        elif tuple(a.shape) != tuple(shape):
            raise ValueError(f"{name}: shape {tuple(a.shape)} != expected {tuple(shape)}")
        out[name] = np.ascontiguousarray(a)
    return out


def from_keras_h5(path, num_classes):
    import h5py                                # not available in the build container (see module docstring)
    convs, tconvs, bns, prelus = [], [], [], []
    with h5py.File(path, "r") as f:
        g = f["model_weights"] if "model_weights" in f else f
        order = [n.decode() if isinstance(n, bytes) else n for n in g.attrs["layer_names"]]
        for lname in order:
            names = [n.decode() if isinstance(n, bytes) else n for n in g[lname].attrs.get("weight_names", [])]
            arrs = [np.asarray(g[lname][n]) for n in names]
            if len(arrs) == 4 and all(a.ndim == 1 for a in arrs):
                bns.append(arrs)               # gamma, beta, moving_mean, moving_variance
            elif len(arrs) >= 1 and arrs[0].ndim == 4:
                (tconvs if "transpose" in lname.lower() else convs).append(arrs[0])
            elif len(arrs) == 1:
                prelus.append(arrs[0].reshape(-1))
    out, ic, it, ib, ip = {}, 0, 0, 0, 0
    spec = W.enet_param_spec(num_classes)
    i = 0
    while i < len(spec):
        name, shape, kind = spec[i]
        if kind == "conv":
            out[name] = np.ascontiguousarray(convs[ic].transpose(3, 2, 0, 1), dtype=np.float32); ic += 1
        elif kind == "tconv":                  # Keras Conv2DTranspose kernel: (kh, kw, out, in)
            out[name] = np.ascontiguousarray(tconvs[it].transpose(3, 2, 0, 1), dtype=np.float32); it += 1
        elif kind == "bn_gamma":
            gamma, beta, mean, var = bns[ib]; ib += 1
            base = name[:-len(".weight")]
            out[base + ".weight"], out[base + ".bias"] = gamma.astype(np.float32), beta.astype(np.float32)
            out[base + ".running_mean"], out[base + ".running_var"] = mean.astype(np.float32), var.astype(np.float32)
            i += 3
        elif kind == "prelu":
            a = prelus[ip]; ip += 1
            out[name] = np.ascontiguousarray(a if a.size == 1 else a.reshape(-1), dtype=np.float32)
        if kind in ("conv", "tconv") and tuple(out[name].shape) != tuple(shape):
            raise ValueError(f"{name}: Keras kernel gives {tuple(out[name].shape)}, expected {tuple(shape)}")
        i += 1
    return {n: out[n] for n, _, _ in spec if n in out}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--classes", type=int, default=15)
    ap.add_argument("--bn-eps", type=float, default=W.BN_EPS)
    a = ap.parse_args()
    if a.src.endswith((".pth", ".pt")):
        import torch
        w = from_state_dict(torch.load(a.src, map_location="cpu"), a.classes)
    elif a.src.endswith(".h5"):
        print("note: the Keras .h5 branch is EXPERIMENTAL -- it has never run (no h5py, no model.h5 in the build "
              "container); it matches tensors by order and shape and fails on the first misfit", file=sys.stderr)
        w = from_keras_h5(a.src, a.classes)
    else:
        sys.exit("expected a .pth/.pt state_dict or a Keras .h5 file")
    hk = int(w["transposed_conv.weight"].shape[-1])
    with open(a.dst, "wb") as f:
        f.write(W.pack_flat(w, a.classes, a.bn_eps, graph=W.graph_rows(), initial_pool=3, head_kernel=hk))
    print(f"wrote {a.dst}: {len(w)} tensors, {a.classes} classes")


if __name__ == "__main__":
    main()
