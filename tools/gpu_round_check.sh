#!/bin/bash
# One GPU-box pass: -m gpu tests, the default bench line, the ncu launch list of a short bench, one `ncu --set full` capture of the soil kernel.
# usage (through gpurun): bash tools/gpu_round_check.sh TAG
tag=${1:-r2}; o=gpurun_out/$tag; mkdir -p $o
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $o/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=30 > $o/gputest.log 2>&1; echo "rc=$?" >> $o/gputest.log
timeout 900 python bench.py > $o/bench.json 2> $o/bench.err; echo "rc=$?" >> $o/bench.err
PROFILE_SITES=151552 PROFILE_DAYS=2 timeout 600 python tools/gpu_profile.py > $o/profile_run.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $o/launches.csv python bench.py --sites 151552 --steps 2 --warmup 3 --no-cpu-baseline --no-probes > $o/ncu_launch.log 2>&1
PROFILE_SITES=151552 PROFILE_DAYS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:soil_chunk -s 1 -c 1 -f -o $o/soil python tools/gpu_profile.py > $o/ncu_full.log 2>&1
# the weather kernel (widening rows f2 / f3): its probe on its own, then one full capture of it
timeout 300 python tools/gpu_weather_probe.py > $o/weather_probe.json 2> $o/weather_probe.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:weather_kernel -s 3 -c 1 -f -o $o/weather python tools/gpu_weather_probe.py --reps 2 > $o/ncu_weather.log 2>&1
ls -la $o
tail -3 $o/gputest.log; cat $o/bench.json | head -c 3000
