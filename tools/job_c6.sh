python tools/conv_probe.py E 0.75,1.5 1 - 10,30,100,300 > gpurun_out/r2_c6_conv_E.txt 2>&1
python tools/conv_probe.py D 1,2 1 - 3,10,30,100 > gpurun_out/r2_c6_conv_D.txt 2>&1
cat gpurun_out/r2_c6_conv_E.txt gpurun_out/r2_c6_conv_D.txt
