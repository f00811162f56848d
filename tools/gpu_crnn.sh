#!/bin/bash
# standard CRNN check on the GPU box: parity tests that touch the CRNN path, then the CRNN-only bench at 512 x 10 s
tag=${1:-crnn}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "crnn or CRNN or encode_detect or deterministic or bench_shape or known_answers or trigger or fp16" > gpurun_out/t_$tag.log 2>&1; tail -3 gpurun_out/t_$tag.log
timeout 300 python bench.py --workload crnn --streams 512 --no-cpu-baseline --no-extras > gpurun_out/b_$tag.json 2> gpurun_out/b_$tag.err
python -c "import json;d=json.load(open('gpurun_out/b_$tag.json'));print(d['ms_per_step'], d['extra']['ms_per_stage'])"
