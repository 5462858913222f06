"""Gaussian point-cloud .ply in the reference's attribute layout (scene/gaussian_model.py:220-232,258-275,
355-395), written / read with numpy only (the reference uses the `plyfile` package):

    x y z nx ny nz f_dc_0..2 f_rest_0..44 opacity scale_0..2 rot_0..3      (all float32, binary little endian)

`f_dc` / `f_rest` are stored channel-major (``transpose(1, 2).flatten(1)``), raw (pre-activation) values.
"""
from __future__ import annotations

import os

import numpy as np
import torch


def attribute_names(n_dc: int, n_rest: int, n_scale: int = 3, n_rot: int = 4):
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(n_dc)]
    names += [f"f_rest_{i}" for i in range(n_rest)]
    names.append("opacity")
    names += [f"scale_{i}" for i in range(n_scale)]
    names += [f"rot_{i}" for i in range(n_rot)]
    return names


def write_gaussian_ply(path, xyz, features_dc, features_rest, opacity, scaling, rotation):
    """Tensors in the model's layout: xyz [P,3], features_dc [P,1,3], features_rest [P,K-1,3], opacity [P,1],
    scaling [P,3], rotation [P,4]."""
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    t = lambda a: a.detach().float().cpu()
    xyz_n = t(xyz).numpy()
    f_dc = t(features_dc).transpose(1, 2).flatten(start_dim=1).contiguous().numpy()
    f_rest = t(features_rest).transpose(1, 2).flatten(start_dim=1).contiguous().numpy()
    cols = np.concatenate((xyz_n, np.zeros_like(xyz_n), f_dc, f_rest, t(opacity).numpy(), t(scaling).numpy(),
                           t(rotation).numpy()), axis=1).astype("<f4")
    names = attribute_names(f_dc.shape[1], f_rest.shape[1], scaling.shape[1], rotation.shape[1])
    assert cols.shape[1] == len(names)
    header = "ply\nformat binary_little_endian 1.0\n" + f"element vertex {cols.shape[0]}\n" + \
        "".join(f"property float {n}\n" for n in names) + "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(np.ascontiguousarray(cols).tobytes())


_PLY_TYPES = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "uchar": "u1", "uint8": "u1",
              "char": "i1", "int8": "i1", "short": "<i2", "int16": "<i2", "ushort": "<u2", "uint16": "<u2",
              "int": "<i4", "int32": "<i4", "uint": "<u4", "uint32": "<u4"}


def read_ply_vertices(path):
    """-> numpy structured array of the first element of a binary little-endian (or ascii) .ply."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a .ply file")
        fmt, count, props, in_first = None, None, [], False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated header")
            tok = line.decode("ascii").split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                if count is None:
                    count, in_first = int(tok[2]), True
                else:
                    in_first = False
            elif tok[0] == "property" and in_first:
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties are not supported in the vertex element")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        dt = np.dtype(props)
        if fmt == "binary_little_endian":
            return np.frombuffer(f.read(count * dt.itemsize), dtype=dt, count=count)
        if fmt == "ascii":
            rows = np.loadtxt(f, max_rows=count, ndmin=2)
            out = np.empty(count, dtype=dt)
            for i, (n, _) in enumerate(props):
                out[n] = rows[:, i]
            return out
        raise ValueError(f"{path}: unsupported format {fmt}")


def read_gaussian_ply(path, max_sh_degree: int):
    """-> dict of float32 CPU tensors in the model's layout (gaussian_model.py:355-395)."""
    v = read_ply_vertices(path)
    names = v.dtype.names
    col = lambda n: np.asarray(v[n], dtype=np.float32)
    xyz = np.stack((col("x"), col("y"), col("z")), axis=1)
    opacity = col("opacity")[:, None]
    f_dc = np.stack((col("f_dc_0"), col("f_dc_1"), col("f_dc_2")), axis=1)[:, :, None]          # [P,3,1]
    key = lambda n: int(n.split("_")[-1])
    rest_names = sorted((n for n in names if n.startswith("f_rest_")), key=key)
    if len(rest_names) != 3 * (max_sh_degree + 1) ** 2 - 3:
        raise ValueError(f"{path}: {len(rest_names)} f_rest attributes do not match SH degree {max_sh_degree}")
    f_rest = np.stack([col(n) for n in rest_names], axis=1).reshape(xyz.shape[0], 3, (max_sh_degree + 1) ** 2 - 1) \
        if rest_names else np.zeros((xyz.shape[0], 3, 0), np.float32)
    scales = np.stack([col(n) for n in sorted((n for n in names if n.startswith("scale_")), key=key)], axis=1)
    rots = np.stack([col(n) for n in sorted((n for n in names if n.startswith("rot")), key=key)], axis=1)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float()
    return {"xyz": T(xyz), "features_dc": T(f_dc).transpose(1, 2).contiguous(),
            "features_rest": T(f_rest).transpose(1, 2).contiguous(), "opacity": T(opacity), "scaling": T(scales),
            "rotation": T(rots)}
