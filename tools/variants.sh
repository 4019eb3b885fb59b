#!/bin/bash
# Kernel experiments: build single-layout variants of libvsr_b200.so (seconds each) so ONE gpurun call can A/B them.
#   tools/variants.sh                       builds build/variants/libvsr_b200_<name>.so for the variants below (cfg2 layout)
#   tools/variants.sh R V K                 same for another layout
# On the GPU box:  for v in build/variants/*.so; do VSR_B200_LIB=$v python tools/quick.py 3 2 2 0 0; done
# Each variant is the default kernel plus -D flags (see "#ifdef VSR_EXP_" in csrc/vsr_gpu.cuh); "base" has none.
set -e
R=${1:-3}; V=${2:-2}; K=${3:-3}
cd "$(dirname "$0")/../vsr-tlaplus_b200/csrc"
OUT=../../build/variants; mkdir -p $OUT
ARCH="-gencode arch=compute_100a,code=sm_100a"
ONLY="-DVSR_ONLY_R=$R -DVSR_ONLY_V=$V -DVSR_ONLY_K=$K"
build() { # name, flags
    g++ -O2 -std=c++17 -fPIC $ONLY $2 -c vsr_host.cpp -o $OUT/vsr_host_$1.o   # the host side shares the flags (fingerprints must agree)
    nvcc $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -diag-suppress 128 $ONLY $2 -c vsr_gpu.cu -o $OUT/vsr_gpu_$1.o
    nvcc $ARCH -shared -Xlinker -Bsymbolic -o $OUT/libvsr_b200_$1.so $OUT/vsr_gpu_$1.o $OUT/vsr_host_$1.o -ldl
    rm -f $OUT/vsr_gpu_$1.o $OUT/vsr_host_$1.o
    echo "built $OUT/libvsr_b200_$1.so"
}
build base ""
build emit_uv "-DVSR_EXP_EMIT_UV"
build home_lowbits "-DVSR_EXP_HOME_LOWBITS"
build prefetch "-DVSR_EXP_PREFETCH"
build par128 "-DVSR_EXP_PAR128"           # parents loaded 16 bytes at a time
build skew "-DVSR_EXP_SKEW"               # scratch rows skewed, not rotated: apply<G> bodies 20 % smaller
build all5 "-DVSR_EXP_EMIT_UV -DVSR_EXP_HOME_LOWBITS -DVSR_EXP_PREFETCH -DVSR_EXP_SKEW -DVSR_EXP_PAR128"
build warps12 "-DVSR_FORCE_WARPS=12"     # 80 registers per thread, 24 warps per SM
build warps8 "-DVSR_FORCE_WARPS=8"       # 128 registers per thread
build qps1 "-DVSR_QPS=1"                 # correctness variant: VSR_B200_LIB=...qps1.so python -m pytest tests/test_gpu_parity.py -k "3-2-2 or deterministic"
build fasthash "-DVSR_EXP_FASTHASH"   # not TLC's fingerprint: measures what FP64 costs, nothing else
