#!/bin/bash
# Development build of the solve kernels into exp/NAME.so (git-ignored, travels to the GPU box):
#     bash profiles/r02/tools/mk.sh NAME [-DFLAG ...]        (add -DCRL_DEV_FAST for ThreeLink / FP64 only)
# Load it with CURIOUSRL_B200_LIB=exp/NAME.so or profiles/r02/tools/ab.py exp/A.so,exp/B.so CASE...
set -e
cd "$(dirname "$0")/../../.."
NAME=$1; shift
mkdir -p exp
C=curiousrl_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -fmad=false -Xcompiler -fPIC"
nvcc $FLAGS "$@" -c -o exp/$NAME.kernels.o $C/ilqr_kernels.cu &
if [ ! -f exp/generic.o ] || [ $C/ilqr_generic.cu -nt exp/generic.o ] || [ $C/ilqr_generic.cuh -nt exp/generic.o ]; then
  nvcc $FLAGS -c -o exp/generic.o $C/ilqr_generic.cu &
fi
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o exp/$NAME.so exp/$NAME.kernels.o exp/generic.o
rm -f exp/$NAME.kernels.o
ls -la exp/$NAME.so
