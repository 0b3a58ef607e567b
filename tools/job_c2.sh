python -m pytest tests/test_fused_gpu.py tests/test_full_size_gpu.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_c2_tests.log
python tools/devbench.py E > gpurun_out/r2_c2_dev_E_new.txt 2>&1
LMC_NO_COL512=1 python tools/devbench.py E > gpurun_out/r2_c2_dev_E_old.txt 2>&1
LMC_COL512_PPC=4 python tools/devbench.py E > gpurun_out/r2_c2_dev_E_ppc4.txt 2>&1
python tools/dbg_case4.py > gpurun_out/dbg4_plain.log 2>&1
compute-sanitizer --tool racecheck python tools/dbg_case4.py > gpurun_out/dbg4_race.log 2>&1
cat gpurun_out/r2_c2_tests.log gpurun_out/r2_c2_dev_E_new.txt gpurun_out/r2_c2_dev_E_old.txt gpurun_out/r2_c2_dev_E_ppc4.txt
tail -12 gpurun_out/dbg4_plain.log; echo ====; tail -12 gpurun_out/dbg4_race.log
