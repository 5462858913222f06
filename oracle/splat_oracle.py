"""numpy/ctypes front-end of oracle/splat_oracle.c (TEST INFRASTRUCTURE ONLY).

`build()` compiles the C restatement with gcc; `Oracle` runs the reference's
rasterizer algorithm on the CPU: forward (preprocess, binning, composite) and
backward.  See the header of splat_oracle.c for what it restates and how it is
pinned.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "splat_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libsplat_oracle.so")

_lib = None


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c11", "-shared", "-fPIC",
           "-o", LIB, SRC, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed building the oracle:\n" + r.stderr)
    return LIB


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        lib.oracle_create.restype = C.c_void_p
        lib.oracle_destroy.argtypes = [C.c_void_p]
        lib.oracle_forward.restype = C.c_int64
        lib.oracle_forward.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int] + \
            [C.c_void_p] * 5 + [C.c_float] + [C.c_void_p] * 5 + [C.c_float, C.c_float] + [C.c_void_p] * 3
        lib.oracle_num_rendered.restype = C.c_int64
        lib.oracle_num_rendered.argtypes = [C.c_void_p]
        lib.oracle_get_geometry.argtypes = [C.c_void_p] * 8
        lib.oracle_get_binning.argtypes = [C.c_void_p] * 4
        lib.oracle_get_image.argtypes = [C.c_void_p] * 3
        lib.oracle_backward.restype = C.c_int
        lib.oracle_backward.argtypes = [C.c_void_p] * 6 + [C.c_float] + [C.c_void_p] * 5 + \
            [C.c_float, C.c_float] + [C.c_void_p] * 12
        lib.oracle_mark_visible.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One forward (+ optional backward) of the reference rasterizer on the CPU."""

    def __init__(self):
        self.lib = _load()
        self.h = C.c_void_p(self.lib.oracle_create())

    def __del__(self):
        try:
            self.lib.oracle_destroy(self.h)
        except Exception:
            pass

    def forward(self, *, bg, W, H, means3D, opacities, viewmatrix, projmatrix, campos, tanfovx, tanfovy,
                sh_degree=0, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, scale_modifier=1.0):
        means3D = _f32(means3D)
        P = means3D.shape[0]
        self.P, self.W, self.H = P, W, H
        shs = _f32(shs); colors_precomp = _f32(colors_precomp)
        M = 0 if shs is None else shs.shape[1]
        self.M = M
        self._in = dict(bg=_f32(bg), means3D=means3D, shs=shs, colors_precomp=colors_precomp,
                        opacities=_f32(opacities), scales=_f32(scales), rotations=_f32(rotations),
                        cov3D_precomp=_f32(cov3D_precomp), view=_f32(viewmatrix).reshape(-1),
                        proj=_f32(projmatrix).reshape(-1), campos=_f32(campos),
                        scale_modifier=float(scale_modifier), tanfovx=float(tanfovx),
                        tanfovy=float(tanfovy), D=int(sh_degree))
        i = self._in
        color = np.zeros((3, H, W), np.float32)
        depth = np.zeros((1, H, W), np.float32)
        radii = np.zeros((P,), np.int32)
        R = self.lib.oracle_forward(self.h, P, i["D"], M, _p(i["bg"]), W, H, _p(means3D), _p(shs),
                                    _p(colors_precomp), _p(i["opacities"]), _p(i["scales"]),
                                    i["scale_modifier"], _p(i["rotations"]), _p(i["cov3D_precomp"]),
                                    _p(i["view"]), _p(i["proj"]), _p(i["campos"]), i["tanfovx"],
                                    i["tanfovy"], _p(color), _p(depth), _p(radii))
        if R < 0:
            raise MemoryError("oracle_forward failed")
        self.R = int(R)
        return color, radii, depth

    def geometry(self):
        P = self.P
        out = dict(depth=np.zeros(P, np.float32), xy=np.zeros((P, 2), np.float32),
                   conic_opacity=np.zeros((P, 4), np.float32), rgb=np.zeros((P, 3), np.float32),
                   cov3D=np.zeros((P, 6), np.float32), clamped=np.zeros((P, 3), np.uint8),
                   tiles_touched=np.zeros(P, np.uint32))
        self.lib.oracle_get_geometry(self.h, _p(out["depth"]), _p(out["xy"]), _p(out["conic_opacity"]),
                                     _p(out["rgb"]), _p(out["cov3D"]), _p(out["clamped"]),
                                     _p(out["tiles_touched"]))
        return out

    def binning(self):
        R = self.R
        gx, gy = (self.W + 15) // 16, (self.H + 15) // 16
        out = dict(keys=np.zeros(R, np.uint64), point_list=np.zeros(R, np.uint32),
                   ranges=np.zeros((gx * gy, 2), np.uint32))
        self.lib.oracle_get_binning(self.h, _p(out["keys"]), _p(out["point_list"]), _p(out["ranges"]))
        return out

    def image_state(self):
        out = dict(final_T=np.zeros(self.H * self.W, np.float32),
                   n_contrib=np.zeros(self.H * self.W, np.uint32))
        self.lib.oracle_get_image(self.h, _p(out["final_T"]), _p(out["n_contrib"]))
        return out

    def backward(self, dL_dcolor, dL_ddepth):
        i = self._in
        P, M = self.P, self.M
        g = dict(means2D=np.zeros((P, 3), np.float32), conic=np.zeros((P, 4), np.float32),
                 opacity=np.zeros((P, 1), np.float32), colors=np.zeros((P, 3), np.float32),
                 depth=np.zeros((P,), np.float32), means3D=np.zeros((P, 3), np.float32),
                 cov3D=np.zeros((P, 6), np.float32), sh=np.zeros((P, M, 3), np.float32),
                 scales=np.zeros((P, 3), np.float32), rotations=np.zeros((P, 4), np.float32))
        dc = _f32(dL_dcolor); dd = _f32(dL_ddepth)
        rc = self.lib.oracle_backward(self.h, _p(i["bg"]), _p(i["means3D"]), _p(i["shs"]),
                                      _p(i["colors_precomp"]), _p(i["scales"]), i["scale_modifier"],
                                      _p(i["rotations"]), _p(i["cov3D_precomp"]), _p(i["view"]),
                                      _p(i["proj"]), _p(i["campos"]), i["tanfovx"], i["tanfovy"],
                                      _p(dc), _p(dd), _p(g["means2D"]), _p(g["conic"]), _p(g["opacity"]),
                                      _p(g["colors"]), _p(g["depth"]), _p(g["means3D"]), _p(g["cov3D"]),
                                      _p(g["sh"]) if M else None, _p(g["scales"]), _p(g["rotations"]))
        if rc != 0:
            raise MemoryError("oracle_backward failed")
        return g


def mark_visible(means3D, viewmatrix):
    means3D = _f32(means3D)
    out = np.zeros(means3D.shape[0], np.uint8)
    _load().oracle_mark_visible(means3D.shape[0], _p(means3D), _p(_f32(viewmatrix).reshape(-1)), _p(out))
    return out.astype(bool)
