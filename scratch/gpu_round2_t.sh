#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 > gpurun_out/r2t_pytest.txt; tail -6 gpurun_out/r2t_pytest.txt
timeout 600 python scratch/time_classes.py > gpurun_out/r2t_classes.txt 2>&1; cat gpurun_out/r2t_classes.txt
timeout 900 python bench.py 2> gpurun_out/bench_r02_h.err | tail -1 > gpurun_out/bench_r02_h.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r02_h.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['parity']['max_abs_dh'], d['parity']['vs_reference']['n_dh_above_1e-4']); print({k:(round(v['ms_per_step'],3), v.get('kernel')) for k,v in d['configs'].items()}); print(d['drop_in']['ms'], d['drop_in']['ms_first_guess_device'], d['drop_in']['same_vectors'], d['drop_in']['same_vectors_first_guess_device'], d['drop_in_ew']['ms'])"
