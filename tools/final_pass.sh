#!/bin/bash
# End-of-round measurement pass on ONE B200 (run through gpurun from the repo root): parity suite, smoke, the bench line and
# its reference arm, the ncu launch list of the bench command, the side tables (schedule sweep, codec / aggregation / latency
# timings, MSM phases).  Everything lands in gpurun_out/<tag>_*; tools/refresh_profiles.py <tag> copies it into profiles/.
# usage: tools/final_pass.sh [tag]
tag=${1:-r02}; o=gpurun_out; mkdir -p $o
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $o/${tag}_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -7 | tee $o/${tag}_smoke.log
python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err; tail -2 $o/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference_arm.json 2>> $o/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 1 > /dev/null 2>&1
python tools/path_sweep.py > $o/${tag}_path_sweep.json 2>/dev/null; tail -3 $o/${tag}_path_sweep.json
python tools/codec_bench.py 2>/dev/null | tail -1 > $o/${tag}_codec_timings.json
python tools/agg_bench.py 2>/dev/null | tail -1 > $o/${tag}_agg_timings.json
python tools/latency_bench.py 2>/dev/null | tail -1 > $o/${tag}_latency.json
(python tools/msm_phases.py 22 0:1 0:8 7:8; python tools/msm_phases.py 20 0:1 0:8 7:8; python tools/msm_phases.py 16 0:1) > $o/${tag}_msm_phases.json 2>/dev/null
head -c 600 $o/${tag}_bench.json; echo
