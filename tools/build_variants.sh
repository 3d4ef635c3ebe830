#!/bin/bash
# Alternative builds of the C-ABI library for kernel tuning on the GPU box (selected with DQ_DECODING_LIB, see _lib.py):
# warp-role splits / queue depths of the env kernel, plus (OLD_REV=<git rev>) an earlier revision's env kernel as a same-box control.
# The Q-network and exchange objects are compiled once and linked into every variant.
set -e
cd "$(dirname "$0")/.."
CS=deepq_decoding_b200/csrc
OUT=build/variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr"
mkdir -p $OUT
[ -f $OUT/dq_qnet.o ] && [ $OUT/dq_qnet.o -nt $CS/dq_qnet.cu ] || $NVCC $FLAGS -c -o $OUT/dq_qnet.o $CS/dq_qnet.cu
[ -f $OUT/dq_comm.o ] && [ $OUT/dq_comm.o -nt $CS/dq_comm.cu ] || $NVCC $FLAGS -c -o $OUT/dq_comm.o $CS/dq_comm.cu
variant() {   # name, extra flags; DQ_VARIANTS="w2g3 q8" restricts the build to the named ones
    name=$1; shift
    if [ -n "$DQ_VARIANTS" ]; then case " $DQ_VARIANTS " in *" $name "*) ;; *) return 0;; esac; fi
    $NVCC $FLAGS "$@" -c -o $OUT/dq_env_$name.o $CS/dq_env.cu
    $NVCC -shared -o $OUT/libdq_$name.so $OUT/dq_env_$name.o $OUT/dq_qnet.o $OUT/dq_comm.o -lcudart
    rm -f $OUT/dq_env_$name.o
    echo "built $OUT/libdq_$name.so"
}
variant w1g4     -DDQ_WRITERS=1 -DDQ_GENS=4
variant w2g3     -DDQ_WRITERS=2 -DDQ_GENS=3
variant w1g3q8   -DDQ_QDEPTH=8
variant d7w3g4   -DDQ_WRITERS_D7=3 -DDQ_GENS_D7=4
variant d7w2g6   -DDQ_WRITERS_D7=2 -DDQ_GENS_D7=6
if [ -n "$OLD_REV" ]; then      # same-box control: the env kernel of an earlier revision
    mkdir -p $OUT/old && git show $OLD_REV:$CS/dq_env.cu > $OUT/old/dq_env.cu && git show $OLD_REV:$CS/dq_lattice.cuh > $OUT/old/dq_lattice.cuh
    git show $OLD_REV:include/dq_decoding.h > $OUT/old/dq_decoding.h
    sed -i 's#"../../include/dq_decoding.h"#"dq_decoding.h"#' $OUT/old/dq_env.cu
    grep -q dq_env_set_max_attempts $OUT/old/dq_env.cu || echo 'extern "C" int dq_env_set_max_attempts(dq_env*, int) { return 0; }' >> $OUT/old/dq_env.cu      # entry points added since
    $NVCC $FLAGS -c -o $OUT/dq_env_old.o $OUT/old/dq_env.cu
    $NVCC -shared -o $OUT/libdq_old.so $OUT/dq_env_old.o $OUT/dq_qnet.o $OUT/dq_comm.o -lcudart
    rm -f $OUT/dq_env_old.o
    echo "built $OUT/libdq_old.so"
fi
