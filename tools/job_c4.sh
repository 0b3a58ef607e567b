python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/r2_c4_tests.log
python tools/devbench.py E > gpurun_out/r2_c4_dev_E.txt 2>&1
LMC_COL512_PPC=2 python tools/devbench.py E > gpurun_out/r2_c4_dev_E_ppc2.txt 2>&1
LMC_P=32 tools/ncu_capture.sh r2_c4_col512 fused_col512 0 1 python tools/one_step.py E mvm_sorted > /dev/null 2>&1
python tools/conv_probe.py E 1.5,4 0.1,1 jacobi > gpurun_out/r2_c4_conv_E.txt 2>&1
python tools/conv_probe.py D 2,4 0.1,1 jacobi > gpurun_out/r2_c4_conv_D.txt 2>&1
cat gpurun_out/r2_c4_tests.log gpurun_out/r2_c4_dev_E.txt gpurun_out/r2_c4_dev_E_ppc2.txt gpurun_out/r2_c4_conv_E.txt gpurun_out/r2_c4_conv_D.txt
