#!/bin/bash
# Kernel-variant experiments: builds mcarray_b200/variants/lib_<tag>.so with extra -D flags for gcc.cu / stft.cu and (on the GPU
# box) benches each through MCAG_LIB_PATH (VAR_FILES="fan_tc ..." picks other sources).   usage: tools/variants.sh build tag="-DFOO -DBAR" ... | tools/variants.sh run workload tag...
cd "$(dirname "$0")/../mcarray_b200/csrc" || exit 1
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ --expt-relaxed-constexpr -Xptxas -v"
mode=$1; shift
if [ "$mode" = build ]; then
  mkdir -p ../variants build
  for spec in "$@"; do
    tag=${spec%%=*}; defs=${spec#*=}
    ( vobjs=""; excl="_v[A-Za-z0-9]*\.o$"
      for f in ${VAR_FILES:-gcc stft}; do
        $NV $defs -c $f.cu -o build/${f}_$tag.o 2> build/${f}_$tag.ptxas.log || { cat build/${f}_$tag.ptxas.log; exit 1; }
        vobjs="$vobjs build/${f}_$tag.o"; excl="$excl\|build/$f\.o"
      done
      objs=$(ls build/*.o | grep -v "$excl")
      $NV -shared -o ../variants/lib_$tag.so $vobjs $objs -lcudart
      grep -A3 "stft_tdoa_kernelILi1024" build/gcc_$tag.ptxas.log | grep Used | sed "s/^/$tag /" ) &
  done
  wait
else
  w=$1; shift
  cd ../..
  for tag in "$@"; do
    MCAG_LIB_PATH=$PWD/mcarray_b200/variants/lib_$tag.so python bench.py --workload $w --no-cpu-baseline --no-e2e --also none --sustain 0 --steps 30 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$tag $w', 'ms/step %.4f' % d['ms_per_step'], {k: round(v,4) for k,v in r['kernels_ms_per_step'].items()})"
  done
fi
