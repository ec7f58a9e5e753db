#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the checker, never shipped, never on the product path.
#
# Compiles the reference's own pointnet2_batch CUDA extension, UNMODIFIED, straight from the
# sources where they lie under /root/reference (read-only), for sm_100a, and "installs" it the
# way `pip install --target` would: outputs go ONLY to oracle/_ref/ (git-ignored, NOT
# gpurun-ignored, so the built module travels to the GPU box).  Nothing under oracle/_ref/ is
# tracked; no reference source is copied into the repository history.
#
#   oracle/_ref/pcdet/ops/pointnet2/pointnet2_batch/pointnet2_batch_cuda.so   <- 9 TUs of src/
#   oracle/_ref/pcdet/ops/pointnet2/pointnet2_batch/{pointnet2_utils,pointnet2_modules,surface_feature}.py
#                                                  <- the reference's python op layer (installed copy)
#   oracle/_ref/pcdet/models/backbones_3d/{IASSD,PAGNet}_backbone.py   <- its callers (installed copy)
#
# Used by: tests/ (-m gpu: reference-vs-oracle-vs-candidate), tests/golden/make_golden.py,
#          bench.py --impl reference.   Recipe = SURVEY.md Appendix B.
set -euo pipefail
REF=${REF_ROOT:-/root/reference}
SRC=$REF/pcdet/ops/pointnet2/pointnet2_batch
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
PKG=$OUT/pcdet/ops/pointnet2/pointnet2_batch
OBJ=$OUT/obj
if [ ! -d "$SRC/src" ]; then echo "build_ref: $SRC not present (GPU box?) - using prebuilt oracle/_ref if any"; exit 0; fi
mkdir -p "$PKG" "$OBJ"
PY=${PYTHON:-python}
TDIR=$($PY -c 'import torch,os;print(os.path.dirname(torch.__file__))')
PI=$($PY -c "import sysconfig;print(sysconfig.get_paths()['include'])")
TI="-I$TDIR/include -I$TDIR/include/torch/csrc/api/include"
DEFS="-DTORCH_EXTENSION_NAME=pointnet2_batch_cuda -DTORCH_API_INCLUDE_EXTENSION_H"
pids=()
for f in ball_query group_points interpolate sampling pointnet2_api; do
  ( [ "$OBJ/$f.o" -nt "$SRC/src/$f.cpp" ] || g++ -std=c++17 -O2 -fPIC -w $TI -I$PI -I/usr/local/cuda/include $DEFS -c "$SRC/src/$f.cpp" -o "$OBJ/$f.o" ) &
  pids+=($!)
done
for f in sampling_gpu ball_query_gpu group_points_gpu interpolate_gpu; do
  ( [ "$OBJ/$f.o" -nt "$SRC/src/$f.cu" ] || nvcc -std=c++17 -O2 -w -Xcompiler -fPIC $TI -I$PI -gencode arch=compute_100a,code=sm_100a \
       --expt-relaxed-constexpr -c "$SRC/src/$f.cu" -o "$OBJ/$f.o" ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
g++ -shared -o "$PKG/pointnet2_batch_cuda.so" "$OBJ"/*.o -L"$TDIR/lib" -lc10 -ltorch -ltorch_cpu -ltorch_python \
    -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,"$TDIR/lib"
# install the reference's python op layer next to its native module (like pip --target would)
install -m 0644 "$SRC/pointnet2_utils.py" "$SRC/pointnet2_modules.py" "$SRC/surface_feature.py" "$PKG/"
# ... and the two backbones that call it (the reference-arm of bench.py runs them unmodified); the
# package __init__ files are EMPTY here so none of pcdet's unrelated dependencies (spconv, SharedArray,
# easydict, ...) are imported.
BB=$OUT/pcdet/models/backbones_3d
mkdir -p "$BB"
install -m 0644 "$REF/pcdet/models/backbones_3d/IASSD_backbone.py" "$REF/pcdet/models/backbones_3d/PAGNet_backbone.py" "$BB/"
for d in "$OUT/pcdet" "$OUT/pcdet/ops" "$OUT/pcdet/ops/pointnet2" "$PKG" "$OUT/pcdet/models" "$BB"; do : > "$d/__init__.py"; done

# ---- the consumer of the path (head + rotated NMS), rebuilt the same way -----------
#   oracle/_ref/pcdet/ops/iou3d_nms/iou3d_nms_cuda.so            <- 4 TUs (incl. the reference's CPU IoU)
#   oracle/_ref/pcdet/ops/roiaware_pool3d/roiaware_pool3d_cuda.so <- 2 TUs (imported by box_utils / the head)
#   + installed copies of iou3d_nms_utils.py, roiaware_pool3d_utils.py, model_nms_utils.py, IASSD_head.py,
#     point_head_template.py and pcdet/utils/{box_coder_utils,box_utils,loss_utils,common_utils}.py
build_ext () {  # name  srcdir  outdir  files...
  local name=$1 sdir=$2 odir=$3; shift 3
  local o=$OBJ/$name; mkdir -p "$o" "$odir"
  local d="-DTORCH_EXTENSION_NAME=$name -DTORCH_API_INCLUDE_EXTENSION_H"
  local ps=()
  for f in "$@"; do
    local b=${f%.*}
    if [ "${f##*.}" = cu ]; then
      ( [ "$o/$b.o" -nt "$sdir/src/$f" ] || nvcc -std=c++17 -O2 -w -Xcompiler -fPIC $TI -I$PI -gencode arch=compute_100a,code=sm_100a \
           --expt-relaxed-constexpr $d -c "$sdir/src/$f" -o "$o/$b.o" ) &
    else
      ( [ "$o/$b.o" -nt "$sdir/src/$f" ] || g++ -std=c++17 -O2 -fPIC -w $TI -I$PI -I/usr/local/cuda/include $d -c "$sdir/src/$f" -o "$o/$b.o" ) &
    fi
    ps+=($!)
  done
  for p in "${ps[@]}"; do wait "$p"; done
  g++ -shared -o "$odir/$name.so" "$o"/*.o -L"$TDIR/lib" -lc10 -ltorch -ltorch_cpu -ltorch_python \
      -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,"$TDIR/lib"
}
IOU=$OUT/pcdet/ops/iou3d_nms
ROI=$OUT/pcdet/ops/roiaware_pool3d
build_ext iou3d_nms_cuda "$REF/pcdet/ops/iou3d_nms" "$IOU" iou3d_cpu.cpp iou3d_nms_api.cpp iou3d_nms.cpp iou3d_nms_kernel.cu
build_ext roiaware_pool3d_cuda "$REF/pcdet/ops/roiaware_pool3d" "$ROI" roiaware_pool3d.cpp roiaware_pool3d_kernel.cu
install -m 0644 "$REF/pcdet/ops/iou3d_nms/iou3d_nms_utils.py" "$IOU/"
install -m 0644 "$REF/pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py" "$ROI/"
UT=$OUT/pcdet/utils; MU=$OUT/pcdet/models/model_utils; DH=$OUT/pcdet/models/dense_heads
mkdir -p "$UT" "$MU" "$DH"
install -m 0644 "$REF/pcdet/utils/box_coder_utils.py" "$REF/pcdet/utils/box_utils.py" "$REF/pcdet/utils/loss_utils.py" \
                "$REF/pcdet/utils/common_utils.py" "$UT/"
install -m 0644 "$REF/pcdet/models/model_utils/model_nms_utils.py" "$MU/"
install -m 0644 "$REF/pcdet/models/dense_heads/IASSD_head.py" "$REF/pcdet/models/dense_heads/point_head_template.py" "$DH/"
for d in "$IOU" "$ROI" "$UT" "$MU" "$DH"; do : > "$d/__init__.py"; done
echo "build_ref: ok -> $PKG"
