#!/bin/bash
# round-2 session AF (1 GPU): row groups for all complex pairings — parity, complex bench chain, complex DMRG through the plugin
OUT=gpurun_out; mkdir -p $OUT
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs
timeout 900 python -m pytest tests/test_contract_gpu.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -1
timeout 300 python tools/synth_sweep.py --ms 2048 --sectors 8 --dist equal --dtypes complex --out $OUT/af_synth_complex.jsonl 2>&1 | tail -1 | cut -c1-300
ITB_ROWGROUPS=0 timeout 300 python tools/synth_sweep.py --ms 2048 --sectors 8 --dist equal --dtypes complex --out $OUT/af_synth_complex_norg.jsonl 2>&1 | tail -1 | cut -c1-300
timeout 900 python -m pytest tests/test_plugin_dmrg.py -m gpu -x -q -k "not trg_per_scale and not two_gpus and not config2" 2>&1 | tail -1
