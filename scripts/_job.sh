set -x
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "afm" 2>&1 | tail -15 > gpurun_out/r2_gputest_n.log
cat gpurun_out/r2_gputest_n.log
timeout 300 python scripts/bench_models.py --only afm_c3 > gpurun_out/r2_afm_v6.json 2>gpurun_out/r2_afm_v6.err
cat gpurun_out/r2_afm_v6.json | cut -c1-260; tail -3 gpurun_out/r2_afm_v6.err
