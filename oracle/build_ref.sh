#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY (oracle/): builds the UNMODIFIED reference extension
# (/root/reference/TCGNN_conv/{TCGNN.cpp,TCGNN_kernel.cu}) from the sources where
# they lie into oracle/_ref/ as a Python module named `TCGNN_ref`.
# The module name is changed with -DTORCH_EXTENSION_NAME only (a compiler flag, the
# sources are untouched) so it can be imported next to the new `TCGNN` module.
# Nothing is copied into the repo; oracle/_ref/ is git-ignored but travels with gpurun.
# Does NOT use the reference's own build system (setup.py); plain nvcc + g++.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${TCGNN_REFERENCE_DIR:-/root/reference}/TCGNN_conv"
OUT="$HERE/_ref"
if [ ! -f "$REF/TCGNN.cpp" ]; then
  echo "[build_ref] reference sources not present at $REF -- skipping (prebuilt files are used if any)"; exit 0
fi
mkdir -p "$OUT"
# The reference's own Python driver, staged UNMODIFIED next to the compiled module (same git-ignored directory, same
# role: it is executed by tests/test_gpu_reference_driver.py on the GPU box -- where /root/reference does not exist
# -- to show that main_tcgnn.py / gnn_conv.py / dataset.py run unchanged on top of the new `TCGNN` module).
DRV="${TCGNN_REFERENCE_DIR:-/root/reference}"
mkdir -p "$OUT/driver"
for f in main_tcgnn.py gnn_conv.py dataset.py config.py; do
  if [ -f "$DRV/$f" ]; then cp -pf "$DRV/$f" "$OUT/driver/$f"; fi
done
( cd "$OUT/driver" && sha256sum *.py > SHA256SUMS 2>/dev/null || true )
PY=python
SUFFIX=$($PY -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))")
TARGET="$OUT/TCGNN_ref$SUFFIX"
if [ -f "$TARGET" ] && [ "$TARGET" -nt "$REF/TCGNN.cpp" ] && [ "$TARGET" -nt "$REF/TCGNN_kernel.cu" ] && [ -z "${FORCE:-}" ]; then
  echo "[build_ref] up to date: $TARGET"; exit 0
fi
TORCH_DIR=$($PY -c "import torch,os;print(os.path.dirname(torch.__file__))" 2>/dev/null)
PYINC=$($PY -c "import sysconfig;print(sysconfig.get_paths()['include'])")
INC="-I$REF -I$TORCH_DIR/include -I$TORCH_DIR/include/torch/csrc/api/include -I$PYINC -I/usr/local/cuda/include"
DEFS="-DTORCH_EXTENSION_NAME=TCGNN_ref -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=1"
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
echo "[build_ref] g++ TCGNN.cpp"
g++ -O2 -fPIC -std=c++17 -w $INC $DEFS -c "$REF/TCGNN.cpp" -o "$TMP/TCGNN.o" &
echo "[build_ref] nvcc TCGNN_kernel.cu (several minutes)"
/usr/local/cuda/bin/nvcc -O2 -std=c++17 -w -Xcompiler -fPIC $INC $DEFS \
   -gencode arch=compute_100,code=sm_100 \
   -D__CUDA_NO_HALF_OPERATORS__ -D__CUDA_NO_HALF_CONVERSIONS__ -D__CUDA_NO_HALF2_OPERATORS__ --expt-relaxed-constexpr \
   -c "$REF/TCGNN_kernel.cu" -o "$TMP/TCGNN_kernel.o"
wait
g++ -shared "$TMP/TCGNN.o" "$TMP/TCGNN_kernel.o" -o "$TARGET" \
   -L"$TORCH_DIR/lib" -L/usr/local/cuda/lib64 -lc10 -ltorch -ltorch_cpu -ltorch_python -lc10_cuda -ltorch_cuda -lcudart \
   -Wl,-rpath,"$TORCH_DIR/lib" -Wl,-rpath,/usr/local/cuda/lib64
echo "[build_ref] built $TARGET"
