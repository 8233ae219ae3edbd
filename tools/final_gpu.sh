#!/bin/bash
# round-end GPU sequence: full GPU suite, the default bench line, the ncu launch list and one --set full capture
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2_final_tests.txt
python bench.py > gpurun_out/r2_final.json 2> gpurun_out/r2_final.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-extra > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_sweep|k_llmse|k_covblend|k_cellterms' -s 4 -c 4 -f -o gpurun_out/r02_c4b python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-extra > gpurun_out/ncu_f.log 2>&1
ncu -i gpurun_out/r02_c4b.ncu-rep --page raw --csv > gpurun_out/r02_c4b_raw.csv 2>/dev/null
cat gpurun_out/r2_final_tests.txt
