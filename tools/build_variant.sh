#!/bin/bash
# tools/build_variant.sh NAME FILE.cu "-DFLAG=.. ..." : A/B build of libsldm_sage.so with one translation unit recompiled
# under extra flags -> build/ab/NAME.so (select it at run time with SLDM_LIB_PATH=build/ab/NAME.so).
set -euo pipefail
name=$1; unit=$2; flags=${3:-}
root=$(cd "$(dirname "$0")/.." && pwd)
cd "$root"
make -j8 sldm_gnn_b200/lib/libsldm_sage.so >/dev/null
mkdir -p build/ab/obj_$name
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr \
     $flags -c sldm_gnn_b200/csrc/$unit.cu -o build/ab/obj_$name/$unit.o
objs=$(ls build/obj/*.o | grep -v "/$unit.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/ab/$name.so $objs build/ab/obj_$name/$unit.o -cudart static
echo "built build/ab/$name.so"
