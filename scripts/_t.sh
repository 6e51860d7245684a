SHORT="python bench.py --mode elementwise --steps 3 --warmup 3 --no-cpu-baseline --no-clustered"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mb_k1_elementwise" -s 4 -c 1 -f -o gpurun_out/r3_k1e $SHORT > gpurun_out/r3_k1e.log 2>&1; echo rc=$?
