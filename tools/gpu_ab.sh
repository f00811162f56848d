#!/bin/bash
# A/B of library variants on the GPU box: tools/gpu_ab.sh <workload> <streams> <tag> [<tag> ...]   ('main' = the in-tree library)
wl=$1; streams=$2; shift 2
for tag in "$@"; do
  lib=$PWD/build/libwwb200_$tag.so
  [ "$tag" == main ] && lib=$PWD/wakeword_detection_b200/libwwb200.so
  WWB200_LIB=$lib timeout 300 python bench.py --workload $wl --streams $streams --no-cpu-baseline --no-extras > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/ab_$tag.json'));print('$tag', d['ms_per_step'], d['extra']['ms_per_stage'])" || tail -3 gpurun_out/ab_$tag.err
done
