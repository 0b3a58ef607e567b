python -m pytest tests/test_fused_gpu.py tests/test_lmc_gpu.py -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r2_c9_tests.log
python tools/devbench.py E --minres 40 > gpurun_out/r2_c9_a.txt 2>&1
LMC_TAIL_CTAS=4 python tools/devbench.py E --minres 40 > gpurun_out/r2_c9_b.txt 2>&1
LMC_TAIL_CTAS=1 python tools/devbench.py E --minres 40 > gpurun_out/r2_c9_c.txt 2>&1
LMC_NO_GRAPH=1 python tools/devbench.py E --minres 40 > gpurun_out/r2_c9_d.txt 2>&1
LMC_NO_OVERLAP=1 python tools/devbench.py E --minres 40 > gpurun_out/r2_c9_e.txt 2>&1
python tools/devbench.py D --minres 100 > gpurun_out/r2_c9_f.txt 2>&1
cat gpurun_out/r2_c9_tests.log; grep -h "minres " gpurun_out/r2_c9_a.txt gpurun_out/r2_c9_b.txt gpurun_out/r2_c9_c.txt gpurun_out/r2_c9_d.txt gpurun_out/r2_c9_e.txt gpurun_out/r2_c9_f.txt
