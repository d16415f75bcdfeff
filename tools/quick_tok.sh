#!/bin/bash
# quick GPU check used while tuning the tokenizer: parity tests that exercise it + the bench line
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python -c "
import json;d=json.load(open('gpurun_out/bench_q.json'));print('value',d['value'],'ms',d['ms_per_step'],'e2e_ms',d['e2e']['ms_per_step'],'tok_ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],d['cpu_baseline']['sample'][-30:])"
