"""Deterministic synthetic parameters keyed by state_dict name.

There is no network for checkpoints, so parity tests and the bench fill BOTH the reference modules
(in the build container, to produce goldens) and the B200 modules with the same values, derived
only from (key name, shape, seed) — independent of construction order and of torch's RNG stream.
"""
import zlib

import numpy as np
import torch


def synthetic_tensor(key, shape, seed=0):
    rng = np.random.default_rng((zlib.crc32(key.encode()) + 7919 * seed) & 0xFFFFFFFF)
    shape = tuple(shape)
    leaf = key.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.long)
    if leaf == "running_var":
        return torch.from_numpy(rng.uniform(0.5, 1.5, size=shape).astype(np.float32))
    if leaf == "running_mean":
        return torch.from_numpy((0.1 * rng.standard_normal(shape)).astype(np.float32))
    is_norm = any(t in key for t in ("norm", "LayerNorm", ".bn.", "loc_layers.0.1", "sem_cls_embed_layer.1")) or \
        (len(shape) == 1 and leaf == "weight")
    if is_norm and leaf == "weight":
        return torch.from_numpy((1.0 + 0.1 * rng.standard_normal(shape)).astype(np.float32))
    if leaf == "bias" or leaf.endswith("_bias"):
        return torch.from_numpy((0.02 * rng.standard_normal(shape)).astype(np.float32))
    if key == "text_features":
        v = rng.standard_normal(shape).astype(np.float32)
        return torch.from_numpy(v / np.linalg.norm(v, axis=-1, keepdims=True))
    fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else int(shape[0])
    std = 1.0 / np.sqrt(max(fan_in, 1))
    return torch.from_numpy((std * rng.standard_normal(shape)).astype(np.float32))


def synthetic_state_dict(module_or_shapes, seed=0):
    """{key: tensor} for every entry of module.state_dict() (or a {key: shape} dict)."""
    if hasattr(module_or_shapes, "state_dict"):
        shapes = {k: tuple(v.shape) for k, v in module_or_shapes.state_dict().items()}
    else:
        shapes = dict(module_or_shapes)
    return {k: synthetic_tensor(k, s, seed) for k, s in shapes.items()}
