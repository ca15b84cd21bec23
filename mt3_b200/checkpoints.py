"""T5X checkpoint reader: what the reference notebook's `restore_from_checkpoint` does with
`gs://mt3/checkpoints/{mt3,ismir2021}` (notebook :247-262, t5x.utils.RestoreCheckpointConfig), without t5x /
flax / tensorstore -- numpy, msgpack and zlib only.

On-disk layout (restated from the T5X `checkpoints.py`, flax `serialization.py` and zarr-v2 specifications;
UNPINNED: no published checkpoint is reachable offline, so the reader is tested against checkpoints written by
`save_t5x_checkpoint` below in the same layout):

  <dir>/checkpoint               msgpack of the train-state dict.  Small arrays are inlined as msgpack ExtType 1 =
                                 packed (shape, dtype name, raw bytes); every large array is replaced by the string
                                 'PLACEHOLDER://<array name>'.
  <dir>/<array name>/.zarray     zarr v2 metadata (shape, chunks, dtype, compressor, order, fill_value)
  <dir>/<array name>/<i>.<j>...  one file per chunk, gzip-compressed C-order bytes (tensorstore's zarr driver)

T5X serialises its train state as {'version': .., 'optimizer': {'target': <params>, 'state': ..}}; the array names (zarr
directories and placeholders) are 'target.' + the parameter path joined with '.', e.g.
'target.decoder.layers_0.mlp.wi_0.kernel' -- NOT prefixed with 'optimizer.'.  That is the layout `save_t5x_checkpoint`
writes by default and the tests read; a top-level 'target' tree is accepted too.  The reader has NOT been run against a
real gs://mt3/checkpoints directory (unreachable offline): treat it as unverified until one can be read.
"""
from __future__ import annotations

import gzip
import itertools
import json
import os
import zlib
from typing import Any, Dict, Mapping

import numpy as np

PLACEHOLDER_PREFIX = "PLACEHOLDER://"


def _ext_hook(code: int, data: bytes):
    import msgpack
    if code == 1:        # flax.serialization._ndarray_to_bytes
        shape, dtype_name, buf = msgpack.unpackb(data, raw=False)
        return np.frombuffer(buf, dtype=np.dtype(dtype_name)).reshape(shape).copy()
    if code == 2:        # native complex
        re, im = msgpack.unpackb(data, raw=False)
        return complex(re, im)
    if code == 3:        # numpy scalar
        shape, dtype_name, buf = msgpack.unpackb(data, raw=False)
        return np.frombuffer(buf, dtype=np.dtype(dtype_name)).reshape(shape)[()]
    return data


def read_zarr_array(path: str) -> np.ndarray:
    """One zarr-v2 array directory -> numpy (gzip / zlib / uncompressed chunks, C or F order, any chunk grid)."""
    with open(os.path.join(path, ".zarray")) as f:
        meta = json.load(f)
    if meta.get("zarr_format", 2) != 2:
        raise ValueError(f"{path}: zarr_format {meta.get('zarr_format')} is not supported (expected 2)")
    shape, chunks = tuple(meta["shape"]), tuple(meta["chunks"])
    dtype = np.dtype(meta["dtype"])
    order = meta.get("order", "C")
    comp = meta.get("compressor")
    sep = meta.get("dimension_separator", ".")
    if meta.get("filters"):
        raise ValueError(f"{path}: zarr filters are not supported")
    fill = meta.get("fill_value")
    out = np.full(shape, 0 if fill is None else fill, dtype=dtype)
    grid = [range(-(-s // c)) for s, c in zip(shape, chunks)] if shape else [range(1)]
    for idx in itertools.product(*grid):
        name = sep.join(str(i) for i in idx) if shape else "0"
        fn = os.path.join(path, name)
        if not os.path.exists(fn):
            continue                                     # missing chunk = fill_value (zarr semantics)
        with open(fn, "rb") as f:
            raw = f.read()
        if comp is not None:
            cid = comp.get("id")
            if cid == "gzip":
                raw = gzip.decompress(raw)
            elif cid == "zlib":
                raw = zlib.decompress(raw)
            else:
                raise ValueError(f"{path}: compressor {cid!r} is not supported (gzip / zlib / none)")
        chunk = np.frombuffer(raw, dtype=dtype).reshape(chunks, order=order)
        sel = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, chunks, shape))
        out[sel] = chunk[tuple(slice(0, s.stop - s.start) for s in sel)]
    return out


def _flatten(tree: Mapping[str, Any], prefix: str = "") -> Dict[str, Any]:
    flat: Dict[str, Any] = {}
    for k, v in tree.items():
        key = f"{prefix}/{k}" if prefix else str(k)
        if isinstance(v, Mapping):
            flat.update(_flatten(v, key))
        else:
            flat[key] = v
    return flat


def load_t5x_checkpoint(path: str) -> Dict[str, np.ndarray]:
    """<dir> (or <dir>/checkpoint) -> {Flax tree path: float32 array}, e.g. 'decoder/layers_0/mlp/wi_0/kernel'."""
    import msgpack
    ckpt_dir = path if os.path.isdir(path) else os.path.dirname(path)
    with open(os.path.join(ckpt_dir, "checkpoint"), "rb") as f:
        state = msgpack.unpackb(f.read(), ext_hook=_ext_hook, raw=False, strict_map_key=False)
    if "optimizer" in state and isinstance(state["optimizer"], Mapping) and "target" in state["optimizer"]:
        target = state["optimizer"]["target"]           # the layout T5X writes
    elif "target" in state:
        target = state["target"]
    else:
        raise ValueError(f"{ckpt_dir}/checkpoint: no 'target' parameter tree (keys: {sorted(state)})")
    params: Dict[str, np.ndarray] = {}
    for key, leaf in _flatten(target).items():
        if isinstance(leaf, str) and leaf.startswith(PLACEHOLDER_PREFIX):
            arr = read_zarr_array(os.path.join(ckpt_dir, leaf[len(PLACEHOLDER_PREFIX):]))
        elif isinstance(leaf, np.ndarray):
            arr = leaf
        else:
            raise ValueError(f"{ckpt_dir}/checkpoint: leaf {key!r} is neither an array nor a placeholder ({type(leaf).__name__})")
        params[key] = np.ascontiguousarray(arr, np.float32)
    return params


def save_t5x_checkpoint(path: str, params: Mapping[str, np.ndarray], step: int = 0, inline_below: int = 4096,
                        max_chunk: int = 512, top_level_target: bool = False) -> None:
    """Write {tree path: array} in the layout described above: {'version', 'optimizer': {'target', 'state'}}, arrays
    with fewer than `inline_below` elements inlined in the msgpack file, the others as gzip zarr arrays named
    'target.<dotted path>' chunked at `max_chunk` per axis.  Used by the tests and to hand-convert weights;
    `top_level_target` puts the tree at state['target'] instead (the variant the reader also accepts)."""
    import msgpack
    os.makedirs(path, exist_ok=True)
    root = "target"

    def pack_array(a: np.ndarray):
        a = np.ascontiguousarray(a)
        return msgpack.ExtType(1, msgpack.packb((list(a.shape), a.dtype.name, a.tobytes()), use_bin_type=True))

    tree: Dict[str, Any] = {}
    for key, arr in params.items():
        arr = np.asarray(arr, np.float32)
        node = tree
        parts = key.split("/")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        if arr.size < inline_below:
            node[parts[-1]] = pack_array(arr)
            continue
        name = root + "." + ".".join(parts)
        node[parts[-1]] = PLACEHOLDER_PREFIX + name
        adir = os.path.join(path, name)
        os.makedirs(adir, exist_ok=True)
        chunks = tuple(min(max_chunk, s) for s in arr.shape)
        with open(os.path.join(adir, ".zarray"), "w") as f:
            json.dump({"zarr_format": 2, "shape": list(arr.shape), "chunks": list(chunks), "dtype": "<f4", "order": "C",
                       "compressor": {"id": "gzip", "level": 1}, "fill_value": None, "filters": None}, f)
        for idx in itertools.product(*[range(-(-s // c)) for s, c in zip(arr.shape, chunks)]):
            block = np.zeros(chunks, np.float32)
            sel = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, chunks, arr.shape))
            block[tuple(slice(0, s.stop - s.start) for s in sel)] = arr[sel]
            with open(os.path.join(adir, ".".join(str(i) for i in idx)), "wb") as f:
                f.write(gzip.compress(block.tobytes(), compresslevel=1))
    state: Dict[str, Any] = {"version": 3}
    if top_level_target:
        state["target"] = tree
        state["state"] = {"step": step}
    else:
        state["optimizer"] = {"target": tree, "state": {"step": step, "param_states": {}}}
    with open(os.path.join(path, "checkpoint"), "wb") as f:
        f.write(msgpack.packb(state, use_bin_type=True))
