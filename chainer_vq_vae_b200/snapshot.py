"""Chainer snapshot (NPZ) import/export -- the on-disk format next to the path
(SURVEY.md section 8f-4).

`chainer.serializers.save_npz(trainer)` writes every parameter under its link path; the model
lives under `updater/model:main/` and generate.py:67-81 loads the sub-trees
`encoder/`, `vq/`, `decoder/ema/` | `decoder/target/` (or `decoder/` without the EMA wrapper)
and `condition_embed/`.  Parameter names inside a sub-tree are Chainer link paths, e.g.
`resnet/7/conv/W` -- identical to this package's module tree with '/' for '.'."""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy
import torch

PREFIX = "updater/model:main/"


def _decoder_prefix(keys, use_ema: bool) -> str:
    has_wrapper = any(k.startswith(PREFIX + "decoder/ema/") for k in keys)
    if not has_wrapper:
        return PREFIX + "decoder/"                        # generate.py:77-79
    return PREFIX + ("decoder/ema/" if use_ema else "decoder/target/")   # generate.py:70-76


def load_chainer_snapshot(path, model, use_ema: bool = True) -> Tuple[List[str], List[str]]:
    """Copy a reference snapshot into `model` (a VAE built by this package).  Returns
    (missing, unexpected) key lists like torch's load_state_dict."""
    npz = numpy.load(path) if not hasattr(path, "files") else path
    keys = list(npz.files)
    dec_prefix = _decoder_prefix(keys, use_ema)
    own: Dict[str, torch.nn.Parameter] = dict(model.named_parameters())
    wrapped = any(n.startswith("decoder.target.") for n in own)
    used, missing = set(), []
    with torch.no_grad():
        for name, p in own.items():
            path_ = name.replace(".", "/")
            if path_.startswith("decoder/"):
                rel = path_[len("decoder/"):]
                if wrapped:                              # decoder.target.* / decoder.ema.*
                    which, rel = rel.split("/", 1)
                    key = (PREFIX + "decoder/" + which + "/" + rel
                           if PREFIX + "decoder/" + which + "/" + rel in npz.files
                           else dec_prefix + rel)
                else:
                    key = dec_prefix + rel
            else:
                key = PREFIX + path_
            if key not in npz.files:
                missing.append(key)
                continue
            arr = npz[key]
            if tuple(arr.shape) != tuple(p.shape):
                raise ValueError(f"{key}: snapshot shape {arr.shape} != model shape {tuple(p.shape)}")
            p.copy_(torch.from_numpy(numpy.ascontiguousarray(arr)).to(p.dtype))
            used.add(key)
    unexpected = [k for k in keys if k.startswith(PREFIX) and k not in used and
                  not k.startswith(PREFIX + "decoder/") ]
    return missing, unexpected


OPT_PREFIX = "updater/optimizer:main/"


def save_chainer_snapshot(path, model, optimizer=None, iteration=None) -> None:
    """Write `model`'s parameters under the reference's key layout.  With `optimizer` (this
    package's Adam) the per-parameter moments are stored the way Chainer's serializer lays out
    an optimizer (`updater/optimizer:main/<param path>/{m,v}`, `.../t`) together with the
    iteration count, so that training can be resumed (`train.py --resume`).  The file is
    written through a file object: `path` is used verbatim (numpy.savez would append `.npz`
    to Chainer's extensionless `snapshot_iter_N` names)."""
    out = {}
    for name, p in model.named_parameters():
        out[PREFIX + name.replace(".", "/")] = p.detach().cpu().numpy()
    if optimizer is not None:
        names = {id(p): n for n, p in model.named_parameters()}
        for p, m, v in zip(optimizer.params, optimizer.m, optimizer.v):
            key = OPT_PREFIX + names[id(p)].replace(".", "/")
            out[key + "/m"] = m.detach().cpu().numpy()
            out[key + "/v"] = v.detach().cpu().numpy()
        out[OPT_PREFIX + "t"] = numpy.asarray(optimizer.t, dtype=numpy.int64)
    if iteration is not None:
        out["updater/iteration"] = numpy.asarray(iteration, dtype=numpy.int64)
    if hasattr(path, "write"):
        numpy.savez(path, **out)
    else:
        with open(path, "wb") as f:
            numpy.savez(f, **out)


def load_optimizer_state(path, model, optimizer) -> int:
    """Restore Adam's moments / step count written by save_chainer_snapshot(optimizer=...).
    Returns the stored iteration (0 if absent)."""
    npz = numpy.load(path) if not hasattr(path, "files") else path
    names = {id(p): n for n, p in model.named_parameters()}
    with torch.no_grad():
        for p, m, v in zip(optimizer.params, optimizer.m, optimizer.v):
            key = OPT_PREFIX + names[id(p)].replace(".", "/")
            m.copy_(torch.from_numpy(numpy.ascontiguousarray(npz[key + "/m"])))
            v.copy_(torch.from_numpy(numpy.ascontiguousarray(npz[key + "/v"])))
    optimizer.t = int(npz[OPT_PREFIX + "t"])
    return int(npz["updater/iteration"]) if "updater/iteration" in npz.files else 0
