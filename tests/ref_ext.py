"""Loader + state decoder for the REAL reference extension (oracle/_ref), test-only.

oracle/_ref/diff_gaussian_rasterization is the unmodified reference CUDA
rasterizer built for sm_100 by oracle/build_ref.sh; it travels to the GPU box
as a prebuilt artefact.  The decoders below re-derive the byte layout of its
three opaque buffers (DGR/cuda_rasterizer/rasterizer_impl.cu:155-194,
rasterizer_impl.h:22-31: each field is 128-byte aligned, in declaration order)
so tests can compare internal state bit for bit.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")

_mod = None


def available() -> bool:
    d = os.path.join(REF_DIR, "diff_gaussian_rasterization")
    return os.path.isdir(d) and any(f.startswith("_C") and f.endswith(".so") for f in os.listdir(d))


def load():
    """import the reference module under the name ``ref_diff_gaussian_rasterization``"""
    global _mod
    if _mod is None:
        if not available():
            raise RuntimeError("oracle/_ref is not built (run oracle/build_ref.sh where /root/reference exists)")
        spec = importlib.util.spec_from_file_location(
            "ref_diff_gaussian_rasterization",
            os.path.join(REF_DIR, "diff_gaussian_rasterization", "__init__.py"),
            submodule_search_locations=[os.path.join(REF_DIR, "diff_gaussian_rasterization")])
        mod = importlib.util.module_from_spec(spec)
        sys.modules["ref_diff_gaussian_rasterization"] = mod
        spec.loader.exec_module(mod)
        _mod = mod
    return _mod


def _carve(buf: torch.Tensor, fields):
    """fields: list of (name, np dtype, count). Returns dict of numpy arrays."""
    raw = buf.detach().cpu().numpy()
    base = buf.data_ptr()
    off = 0
    out = {}
    for name, dt, count in fields:
        a = ((base + off + 127) & ~127) - base
        nbytes = np.dtype(dt).itemsize * count
        if name is not None:
            out[name] = raw[a:a + nbytes].view(dt).copy()
        off = a + nbytes
    return out


def decode_geom(buf, P):
    return _carve(buf, [("depths", np.float32, P), ("clamped", np.uint8, 3 * P),
                        ("internal_radii", np.int32, P), ("means2D", np.float32, 2 * P),
                        ("cov3D", np.float32, 6 * P), ("conic_opacity", np.float32, 4 * P),
                        ("rgb", np.float32, 3 * P), ("tiles_touched", np.uint32, P)])


def decode_binning(buf, R):
    return _carve(buf, [("point_list", np.uint32, R), ("point_list_unsorted", np.uint32, R),
                        ("point_list_keys", np.uint64, R)])


def decode_image(buf, N):
    return _carve(buf, [("accum_alpha", np.float32, N), ("n_contrib", np.uint32, N),
                        ("ranges", np.uint32, 2 * N)])


# ---- the reference's PyTorch HexPlane / deformation module (oracle/_ref/s3g_ref) ----
def deform_available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "s3g_ref", "scene", "deformation.py"))


def load_ref_deform():
    """(deform_network class, eval_sh) of the UNMODIFIED reference, imported from the git-ignored
    install dir; `tkinter` (scene/deformation.py:5) and the scene/utils package __init__s are
    stand-ins because the path never touches what they would import."""
    import types
    base = os.path.join(REF_DIR, "s3g_ref")
    sys.modules.setdefault("tkinter", types.ModuleType("tkinter")).W = None
    for name in ("scene", "utils"):
        if name not in sys.modules or not getattr(sys.modules[name], "__path__", None) or \
                base not in str(sys.modules[name].__path__):
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(base, name)]
            sys.modules[name] = m
    from scene.deformation import deform_network
    from utils.sh_utils import eval_sh
    return deform_network, eval_sh


def ref_deform_args(resolution, multires, **flags):
    from argparse import Namespace
    d = dict(net_width=64, timebase_pe=4, defor_depth=1, posebase_pe=10, scale_rotation_pe=2, opacity_pe=2,
             timenet_width=64, timenet_output=32, bounds=1.6, plane_tv_weight=0.0001, time_smoothness_weight=0.01,
             l1_time_planes=0.0001,
             kplanes_config={"grid_dimensions": 2, "input_coordinate_dim": 4, "output_coordinate_dim": 32,
                             "resolution": list(resolution)},
             multires=list(multires), no_dx=False, no_grid=False, no_ds=True, no_dr=True, no_do=True, no_dshs=False,
             feat_head=True, empty_voxel=False, grid_pe=0, static_mlp=False, apply_rotation=False)
    d.update(flags)
    return Namespace(**d)


# ---- the reference's loss functions (oracle/_ref/s3g_ref/utils/loss_utils.py) ----
def loss_utils_available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "s3g_ref", "utils", "loss_utils.py"))


def load_ref_loss_utils():
    """utils/loss_utils.py of the UNMODIFIED reference (l1_loss, ssim, compute_depth)."""
    import types
    base = os.path.join(REF_DIR, "s3g_ref")
    if "utils" not in sys.modules or base not in str(getattr(sys.modules["utils"], "__path__", "")):
        m = types.ModuleType("utils")
        m.__path__ = [os.path.join(base, "utils")]
        sys.modules["utils"] = m
    from utils import loss_utils
    return loss_utils


# ---- the reference's GaussianModel methods (oracle/_ref/s3g_ref/scene/gaussian_model.py) ----
def gaussian_model_available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "s3g_ref", "scene", "gaussian_model.py")) and \
        os.path.isfile(os.path.join(REF_DIR, "s3g_ref", "utils", "general_utils.py"))


def load_ref_gaussian_model_class():
    """A class carrying the UNMODIFIED method bodies of the reference's GaussianModel that the training loop
    uses for densification (properties, densify*, prune*, *_optimizer, reset_opacity, add_densification_stats).
    The module itself cannot be imported (simple_knn, open3d, plyfile): the FunctionDefs are taken from the
    source with ast and compiled as they are; build_rotation / inverse_sigmoid come the same way from
    utils/general_utils.py."""
    import ast
    import torch
    base = os.path.join(REF_DIR, "s3g_ref")
    ns = {"torch": torch, "nn": torch.nn, "np": np}

    def grab(path, names, into):
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):
            if isinstance(node, ast.FunctionDef) and node.name in names:
                exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), into)
    grab(os.path.join(base, "utils", "general_utils.py"), {"build_rotation", "inverse_sigmoid"}, ns)
    methods = {"get_scaling", "get_rotation", "get_xyz", "get_features", "get_opacity", "densify", "densify_and_clone",
               "densify_and_split", "densification_postfix", "cat_tensors_to_optimizer", "prune_points",
               "_prune_optimizer", "prune", "add_densification_stats", "reset_opacity", "replace_tensor_to_optimizer"}
    cls_ns = dict(ns)
    grab(os.path.join(base, "scene", "gaussian_model.py"), methods, cls_ns)

    class RefGaussianModel:
        scaling_activation = staticmethod(torch.exp)
        scaling_inverse_activation = staticmethod(torch.log)
        opacity_activation = staticmethod(torch.sigmoid)
        rotation_activation = staticmethod(torch.nn.functional.normalize)
    for m in methods:
        setattr(RefGaussianModel, m, cls_ns[m])
    # the extracted functions look their globals up in cls_ns (torch, nn, build_rotation, inverse_sigmoid)
    return RefGaussianModel


def load_ref_compute_regulation():
    """-> f(reference deform_network, time_smoothness_weight, l1_time_planes_weight, plane_tv_weight): the UNMODIFIED
    GaussianModel.compute_regulation + helpers (scene/gaussian_model.py:710-749) and compute_plane_smoothness
    (scene/regulation.py:22-28), extracted with ast (the modules import matplotlib / open3d / simple_knn)."""
    import ast
    import types
    import torch
    base = os.path.join(REF_DIR, "s3g_ref", "scene")
    ns = {"torch": torch}

    def grab(path, names):
        for node in ast.walk(ast.parse(open(path).read())):
            if isinstance(node, ast.FunctionDef) and node.name in names:
                exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    grab(os.path.join(base, "regulation.py"), {"compute_plane_smoothness"})
    names = ("_plane_regulation", "_time_regulation", "_l1_regulation", "compute_regulation")
    grab(os.path.join(base, "gaussian_model.py"), set(names))

    def run(net, time_smoothness_weight, l1_time_planes_weight, plane_tv_weight):
        model = types.SimpleNamespace(_deformation=net)
        for m in names[:3]:
            setattr(model, m, types.MethodType(ns[m], model))
        return ns["compute_regulation"](model, time_smoothness_weight, l1_time_planes_weight, plane_tv_weight)
    return run


def regulation_available() -> bool:
    return gaussian_model_available() and os.path.isfile(os.path.join(REF_DIR, "s3g_ref", "scene", "regulation.py"))


# ---- the reference's simple-knn extension (oracle/_ref/simple_knn) ----
def simple_knn_available() -> bool:
    import glob
    return bool(glob.glob(os.path.join(REF_DIR, "simple_knn", "_C*.so")))


def load_ref_simple_knn():
    """distCUDA2 of the UNMODIFIED reference extension."""
    import importlib.util
    import glob
    import torch  # noqa: F401  (the extension links against libtorch)
    so = glob.glob(os.path.join(REF_DIR, "simple_knn", "_C*.so"))[0]
    spec = importlib.util.spec_from_file_location("simple_knn._C", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.distCUDA2
