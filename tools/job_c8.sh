python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2_c8_tests.log
python tools/devbench.py E --minres 40 > gpurun_out/r2_c8_dev_E.txt 2>&1
LMC_NO_OVERLAP=1 python tools/devbench.py E --minres 40 > gpurun_out/r2_c8_dev_E_noov.txt 2>&1
python tools/devbench.py D --minres 100 > gpurun_out/r2_c8_dev_D.txt 2>&1
LMC_NO_OVERLAP=1 python tools/devbench.py D --minres 100 > gpurun_out/r2_c8_dev_D_noov.txt 2>&1
python tools/devbench.py B --minres 200 > gpurun_out/r2_c8_dev_B.txt 2>&1
cat gpurun_out/r2_c8_tests.log; grep -h minres gpurun_out/r2_c8_dev_E.txt gpurun_out/r2_c8_dev_E_noov.txt gpurun_out/r2_c8_dev_D.txt gpurun_out/r2_c8_dev_D_noov.txt gpurun_out/r2_c8_dev_B.txt
