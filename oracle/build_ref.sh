#!/usr/bin/env bash
# Build the UNMODIFIED reference (nnanhuang/S3Gaussian) into oracle/_ref/ so the
# real reference can be used as (a) the parity checker on the GPU box and
# (b) the reference arm / CPU baseline of bench.py.
#
# Nothing here is product code and nothing under oracle/_ref/ is committed
# (.gitignore lists it; it still travels to the GPU box with gpurun).
# Reference sources are compiled/copied from where they lie under
# /root/reference; we do not run the reference's own build system beyond its
# setup.py compile step on a scratch copy (the tree is read-only), and the only
# deviation from "as is" is forcing <cstdint> in (gcc 13 needs it for
# rasterizer_impl.h's std::uintptr_t; SURVEY.md header table).
#
#   oracle/_ref/diff_gaussian_rasterization/{__init__.py,_C*.so}   reference CUDA rasterizer, sm_100
#   oracle/_ref/simple_knn/_C*.so                                  reference simple-knn (distCUDA2), sm_100
#   oracle/_ref/s3g_ref/scene/{hexplane,deformation,grid}.py       reference HexPlane + decoder (PyTorch)
#   oracle/_ref/s3g_ref/scene/gaussian_model.py                    densify / prune methods (extracted with ast by the tests)
#   oracle/_ref/s3g_ref/utils/{graphics_utils,sh_utils,loss_utils,general_utils}.py
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "build_ref: $REF not present (GPU box?) - using prebuilt $OUT if any"; exit 0
fi
DGR="$REF/submodules/depth-diff-gaussian-rasterization"
mkdir -p "$OUT/diff_gaussian_rasterization" "$OUT/s3g_ref/scene" "$OUT/s3g_ref/utils"

# --- python half (pure copies into the git-ignored install dir) -------------
cp -f "$REF/scene/hexplane.py" "$REF/scene/deformation.py" "$REF/scene/grid.py" "$REF/scene/gaussian_model.py" "$REF/scene/regulation.py" "$OUT/s3g_ref/scene/"
cp -f "$REF/utils/graphics_utils.py" "$REF/utils/sh_utils.py" "$REF/utils/loss_utils.py" "$REF/utils/general_utils.py" "$OUT/s3g_ref/utils/"
cp -f "$REF/arguments/__init__.py" "$OUT/s3g_ref/arguments_init.py"
cp -f "$DGR/diff_gaussian_rasterization/__init__.py" "$OUT/diff_gaussian_rasterization/__init__.py"

# --- simple-knn (distCUDA2): reference initialiser of the scales, checker for csrc/knn.cuh -------------
if ! ls "$OUT"/simple_knn/_C*.so >/dev/null 2>&1 || [ -n "${FORCE:-}" ]; then
  KTMP="$(mktemp -d /tmp/sknn_build.XXXXXX)"
  cp -r "$REF/submodules/simple-knn/." "$KTMP/"
  chmod -R u+w "$KTMP"
  # only deviation from "as is": <cfloat>/<cstdint> forced in (FLT_MAX, gcc 13)
  ( cd "$KTMP" && NVCC_PREPEND_FLAGS="-include cfloat -include cstdint" TORCH_CUDA_ARCH_LIST="10.0" MAX_JOBS=8 \
      python setup.py build_ext --inplace >"$KTMP/build.log" 2>&1 ) || { tail -40 "$KTMP/build.log"; exit 1; }
  mkdir -p "$OUT/simple_knn"
  cp "$KTMP"/simple_knn/_C*.so "$OUT/simple_knn/"
  rm -rf "$KTMP"
  echo "build_ref: simple_knn ok"
fi

# --- CUDA half ----------------------------------------------------------------
if ls "$OUT"/diff_gaussian_rasterization/_C*.so >/dev/null 2>&1 && [ -z "${FORCE:-}" ]; then
  echo "build_ref: reference extension already built"; exit 0
fi
TMP="$(mktemp -d /tmp/dgr_build.XXXXXX)"
cp -r "$DGR/." "$TMP/"
chmod -R u+w "$TMP"
( cd "$TMP" && NVCC_PREPEND_FLAGS="-include cstdint" TORCH_CUDA_ARCH_LIST="10.0" MAX_JOBS=8 \
    python setup.py build_ext --inplace >"$TMP/build.log" 2>&1 ) || { tail -40 "$TMP/build.log"; exit 1; }
cp "$TMP"/diff_gaussian_rasterization/_C*.so "$OUT/diff_gaussian_rasterization/"
rm -rf "$TMP"
echo "build_ref: ok -> $OUT"
