python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2_c3_tests.log
LMC_COL512_MINB=1 python tools/devbench.py E > gpurun_out/r2_c3_dev_E_minb1.txt 2>&1
LMC_P=32 tools/ncu_capture.sh r2_c3_col512 fused_col512 0 1 python tools/one_step.py E mvm_sorted > /dev/null 2>&1
python tools/dbg_case4.py > gpurun_out/dbg4_plain.log 2>&1
compute-sanitizer --tool racecheck python tools/dbg_case4.py > gpurun_out/dbg4_race.log 2>&1
python tools/devbench.py B --minres 200 > gpurun_out/r2_c3_dev_B.txt 2>&1
LMC_NO_GRAPH=1 python tools/devbench.py B --minres 200 > gpurun_out/r2_c3_dev_B_nograph.txt 2>&1
python tools/devbench.py E --minres 20 > gpurun_out/r2_c3_dev_E_minres.txt 2>&1
cat gpurun_out/r2_c3_tests.log gpurun_out/r2_c3_dev_E_minb1.txt gpurun_out/r2_c3_dev_B.txt gpurun_out/r2_c3_dev_B_nograph.txt gpurun_out/r2_c3_dev_E_minres.txt
tail -14 gpurun_out/dbg4_plain.log; echo ====; tail -14 gpurun_out/dbg4_race.log
