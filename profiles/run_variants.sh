#!/bin/bash
# usage: profiles/run_variants.sh tag variant1 variant2 ...   (variant "" = shipped library); logs under gpurun_out/
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  lib=hypernerf_torch_b200/libhypernerf_b200${v:+_$v}.so
  [ "$v" = base ] && lib=hypernerf_torch_b200/libhypernerf_b200.so
  echo "=== $v ($lib)" >> gpurun_out/${tag}_kt.log
  HN_LIB=$PWD/$lib python profiles/kernel_times.py 2>&1 | tail -4 >> gpurun_out/${tag}_kt.log
  if [ -f hypernerf_torch_b200/libhypernerf_b200_${v}_t.so ]; then
    echo "=== $v" >> gpurun_out/${tag}_roles.log
    HN_LIB=$PWD/hypernerf_torch_b200/libhypernerf_b200_${v}_t.so python profiles/role_timing.py >> gpurun_out/${tag}_roles.log 2>&1
  fi
  HN_LIB=$PWD/$lib python bench.py --steps 5 --warmup 3 --no-render --no-static --no-cpu-baseline > gpurun_out/${tag}_bench_$v.json 2> gpurun_out/${tag}_bench_$v.err
done
