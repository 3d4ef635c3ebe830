#!/bin/bash
# Alternative builds of the C-ABI library for kernel tuning on the GPU box (selected with DQ_DECODING_LIB, see _lib.py):
# tile shapes / occupancy targets of the env kernel, plus the previous round's env kernel as a same-box control.
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
variant() {   # name, extra flags; DQ_VARIANTS="so dfso" restricts the build to the named ones
    name=$1; shift
    if [ -n "$DQ_VARIANTS" ]; then case " $DQ_VARIANTS " in *" $name "*) ;; *) return 0;; esac; fi
    $NVCC $FLAGS "$@" -c -o $OUT/dq_env_$name.o $CS/dq_env.cu
    $NVCC -shared -o $OUT/libdq_$name.so $OUT/dq_env_$name.o $OUT/dq_qnet.o $OUT/dq_comm.o -lcudart
    rm -f $OUT/dq_env_$name.o
    echo "built $OUT/libdq_$name.so"
}
variant df         -DDQ_DEFER=1
variant so         -DDQ_STREAM_OBS=1
variant dfso       -DDQ_DEFER=1 -DDQ_STREAM_OBS=1
variant df2        -DDQ_DEFER=2
variant df2so      -DDQ_DEFER=2 -DDQ_STREAM_OBS=1
variant df2solk    -DDQ_DEFER=2 -DDQ_STREAM_OBS=1 -DDQ_LUT_KEEP=1
variant df2sot160  -DDQ_DEFER=2 -DDQ_STREAM_OBS=1 -DDQ_THREADS=160 -DDQ_MIN_BLOCKS=7      # five warps: one more for phase B and the helpers; 56 registers, ~100 B of spills
variant lk         -DDQ_LUT_KEEP=1
variant df2sopf1   -DDQ_DEFER=2 -DDQ_STREAM_OBS=1 -DDQ_PREFETCH=1 -DDQ_REFILL=1
variant df2sopf2   -DDQ_DEFER=2 -DDQ_STREAM_OBS=1 -DDQ_PREFETCH=1 -DDQ_REFILL=2
variant dfsopf1    -DDQ_DEFER=1 -DDQ_STREAM_OBS=1 -DDQ_PREFETCH=1 -DDQ_REFILL=1
variant bb2so      -DDQ_BATCHB=2 -DDQ_STREAM_OBS=1
variant bb2        -DDQ_BATCHB=2
variant bb         -DDQ_BATCHB=1
variant bbmb8      -DDQ_BATCHB=1 -DDQ_MIN_BLOCKS=8
variant bbe32      -DDQ_BATCHB=1 -DDQ_EPC=32 -DDQ_THREADS=256 -DDQ_MIN_BLOCKS=4 -DDQ_STREAM_OBS=1     # fewest instructions per lattice-step of all builds
variant bbso       -DDQ_BATCHB=1 -DDQ_STREAM_OBS=1
variant bbsomi     -DDQ_BATCHB=1 -DDQ_STREAM_OBS=1 -DDQ_MIRROR=1      # + frame / counters / action boards read from shared memory in phase A
variant bb2somi    -DDQ_BATCHB=2 -DDQ_STREAM_OBS=1 -DDQ_MIRROR=1
variant pf1        -DDQ_PREFETCH=1 -DDQ_REFILL=1
variant pf2        -DDQ_PREFETCH=1
variant pf3        -DDQ_PREFETCH=1 -DDQ_REFILL=3
variant pf2mb8     -DDQ_PREFETCH=1 -DDQ_MIN_BLOCKS=8
variant mb8        -DDQ_MIN_BLOCKS=8
variant mb10       -DDQ_MIN_BLOCKS=10
variant e8t64mb14  -DDQ_EPC=8  -DDQ_THREADS=64  -DDQ_MIN_BLOCKS=14
variant e16t96     -DDQ_EPC=16 -DDQ_THREADS=96  -DDQ_MIN_BLOCKS=7
variant bb2t96     -DDQ_BATCHB=2 -DDQ_THREADS=96 -DDQ_MIN_BLOCKS=7
variant e32t256mb4 -DDQ_EPC=32 -DDQ_THREADS=256 -DDQ_MIN_BLOCKS=4
variant e16t256mb4 -DDQ_EPC=16 -DDQ_THREADS=256 -DDQ_MIN_BLOCKS=4
if [ -n "$OLD_REV" ]; then    # control: the env kernel of an earlier commit, built in a scratch copy
    T=$(mktemp -d)
    mkdir -p $T/deepq_decoding_b200/csrc $T/include
    for f in dq_env.cu dq_lattice.cuh dq_ptx.cuh; do git show $OLD_REV:$CS/$f > $T/$CS/$f; done
    git show $OLD_REV:include/dq_decoding.h > $T/include/dq_decoding.h
    $NVCC $FLAGS -c -o $OUT/dq_env_old.o $T/$CS/dq_env.cu
    $NVCC -shared -o $OUT/libdq_old.so $OUT/dq_env_old.o $OUT/dq_qnet.o $OUT/dq_comm.o -lcudart
    rm -rf $T $OUT/dq_env_old.o
    echo "built $OUT/libdq_old.so ($OLD_REV)"
fi
