#!/bin/bash
# Builds libmdf_b200.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libmdf_b200.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=default --use_fast_math=false -Xptxas -v"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v ${MDF_EXTRA_NVCC:-}"
mkdir -p ../../build/obj
objs=()
for f in *.cu; do
  o=../../build/obj/${f%.cu}.o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find . -name '*.cuh' -newer "$o")" ] || [ ../../include/mdf_b200.h -nt "$o" ]; then
    echo "nvcc $f"
    $NVCC $FLAGS -c "$f" -o "$o" 2> "../../build/obj/${f%.cu}.ptxas.log" || { cat "../../build/obj/${f%.cu}.ptxas.log"; exit 1; }
  fi
  objs+=("$o")
done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a "${objs[@]}" -o $OUT -lcuda
echo "built $OUT"
# CPython glue (lists of str / ndarray -> pointer arrays for the ragged entry points); no CUDA in it
PY=${PYTHON:-python3}
PYINC=$($PY -c "import sysconfig; print(sysconfig.get_paths()['include'])")
PYEXT=$($PY -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")
HOSTSO=../_mdf_pyhost$PYEXT
if [ ! -f "$HOSTSO" ] || [ pyhost.c -nt "$HOSTSO" ]; then
  ${CC:-gcc} -O2 -Wall -shared -fPIC -I"$PYINC" pyhost.c -o "$HOSTSO"
  echo "built $HOSTSO"
fi
