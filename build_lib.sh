#!/bin/bash
# Builds audio_ident_b200/libaudioident_b200.so for sm_100a (called by __graft_entry__.build()).
set -e
cd "$(dirname "$0")/audio_ident_b200/csrc"
NVCC=${NVCC:-nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O2"
mkdir -p ../../build/obj
objs=""
pids=""
for f in stft peaks scan hasher synth index match exchange dedup resample engine; do
  [ -f $f.cu ] || continue
  o=../../build/obj/$f.o
  if [ ! -f $o ] || [ $f.cu -nt $o ] || [ common.cuh -nt $o ] || [ engine.h -nt $o ] || [ index.h -nt $o ] || [ ../../include/aid_params.h -nt $o ] || [ ../../include/audio_ident_b200.h -nt $o ]; then
    $NVCC $FLAGS -c $f.cu -o $o &
    pids="$pids $!"
  fi
  objs="$objs $o"
done
for p in $pids; do wait $p; done
$NVCC -Wno-deprecated-gpu-targets -shared -o ../libaudioident_b200.so $objs -lcudart
echo built audio_ident_b200/libaudioident_b200.so
