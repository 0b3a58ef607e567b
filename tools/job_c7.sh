python -m pytest tests/test_fused_gpu.py tests/test_full_size_gpu.py -q -m gpu -x 2>&1 | tail -25 > gpurun_out/r2_c7_tests.log
python tools/devbench.py E > gpurun_out/r2_c7_dev_E.txt 2>&1
LMC_NO_ROWS512=1 python tools/devbench.py E > gpurun_out/r2_c7_dev_E_norows.txt 2>&1
cat gpurun_out/r2_c7_tests.log gpurun_out/r2_c7_dev_E.txt gpurun_out/r2_c7_dev_E_norows.txt
