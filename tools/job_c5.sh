python tools/devbench.py E > gpurun_out/r2_c5_dev_E.txt 2>&1
python tools/conv_probe.py E 0.75,1.5 3,10,30 > gpurun_out/r2_c5_conv_E.txt 2>&1
python tools/conv_probe.py D 1,2 3,10,30 > gpurun_out/r2_c5_conv_D.txt 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_c5_bench_E.json 2> gpurun_out/r2_c5_bench_E.err
cat gpurun_out/r2_c5_dev_E.txt gpurun_out/r2_c5_conv_E.txt gpurun_out/r2_c5_conv_D.txt; tail -5 gpurun_out/r2_c5_bench_E.err; head -c 3000 gpurun_out/r2_c5_bench_E.json
