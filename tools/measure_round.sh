#!/bin/bash
# Round-end measurement pass on the GPU box: tests, ncu captures, pipeline profiles, sanitizers -> gpurun_out/fin/.
# (bench.py lines are taken by a second call, after bench.py's NCU_DRAM table has been updated from the captures.)
mkdir -p gpurun_out/fin
O=gpurun_out/fin
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --leaves 100000 --no-cpu-baseline --parity-sample 0 > $O/launches_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:encode_tc_kernel -s 2 -c 1 -o $O/prof_encode python tools/profile_kernels.py --leaves 59200 > $O/prof_encode.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:decode_tc_kernel -s 2 -c 1 -o $O/prof_decode python tools/profile_kernels.py --leaves 59200 > $O/prof_decode.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:encode_tc128 -c 2 -o $O/prof_vec3_encode python tools/time_vec3_encode.py 4144 1 > $O/prof_vec3_encode.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:decode_tc128 -s 1 -c 1 -o $O/prof_vec3_decode python tools/profile_kernels.py --vec3-decode --leaves 29600 > $O/prof_vec3_decode.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_vec3.csv python bench.py --workload vec3 --steps 1 --warmup 3 --leaves 41440 --no-cpu-baseline --parity-sample 0 > $O/launches_vec3_bench.log 2>&1
timeout 300 python tools/enc_tc_prof.py > $O/encode_pipeline.txt 2>&1
timeout 300 python tools/check_vec3_encode.py 16384 > $O/vec3_encode_check.txt 2>&1
for t in memcheck synccheck; do
  (echo "# compute-sanitizer --tool $t python tools/sanitizer_run.py 301 (final build)"; timeout 600 compute-sanitizer --tool $t python tools/sanitizer_run.py 301 2>&1 | grep -v "Host Frame\|^=========         in\|^=========     at" | tail -40) > $O/sanitizer_$t.txt
done
(echo "# compute-sanitizer --tool racecheck python tools/sanitizer_run.py 151 (final build)"; timeout 600 compute-sanitizer --tool racecheck python tools/sanitizer_run.py 151 2>&1 | grep -v "Host Frame\|^=========         in\|^=========     at" | tail -60) > $O/sanitizer_racecheck.txt
tail -3 $O/sanitizer_memcheck.txt $O/sanitizer_synccheck.txt $O/sanitizer_racecheck.txt | cat
ls -la $O
