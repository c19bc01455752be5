#!/bin/bash
# Round evidence on a GPU box: tests, bench lines, timelines, ncu launch list + full captures, sanitizer.
# Everything lands in gpurun_out/ (copied to profiles/ by hand after reading).  Usage: bash tools/evidence.sh [tag]
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/${TAG}_tests.log
timeout 400 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 200 --warmup 3 2>/dev/null | tail -1 > $O/${TAG}_bench_ref.json
timeout 300 python bench.py --config c2 2>/dev/null | tail -1 > $O/${TAG}_bench_c2.json
timeout 300 python bench.py --config c3 2>/dev/null | tail -1 > $O/${TAG}_bench_c3.json
timeout 120 python tools/step_trace.py packed=3 2>&1 | grep -A6 "graph of 1" > $O/${TAG}_step_trace.txt
timeout 120 python tools/encoder_time.py > $O/${TAG}_encoder_time.txt 2>&1
# launch list of the decode bench: our kernels only (namespace sfb), 12 eager steps + the per-episode projection
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:sfb -c 80 --csv \
  --log-file $O/${TAG}_launches_decode.csv python bench.py --profile-steps 12 > $O/${TAG}_ncu_list.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:sfb -c 40 --csv \
  --log-file $O/${TAG}_launches_encoder.csv python tools/encoder_time.py > /dev/null 2>&1
# one full capture of the step kernel and of the persistent encoder kernel
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:step_kernel -s 4 -c 1 -f \
  -o $O/${TAG}_step_kernel python bench.py --profile-steps 8 > $O/${TAG}_ncu_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:encoder_persist -s 1 -c 1 -f \
  -o $O/${TAG}_encoder_persist python tools/encoder_time.py >> $O/${TAG}_ncu_full.log 2>&1
# memcheck over the fast paths (packed decode incl. the one-launch step, encoder, training backward)
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_packed.py tests/test_gpu_train.py -q \
  -k "carry or graph or tail or encoder or backward" 2>&1 | tail -12 > $O/${TAG}_sanitizer.log
echo done
