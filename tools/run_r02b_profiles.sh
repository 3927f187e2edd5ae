# round-2 (final binary) ncu evidence of the headline launch: every capture follows a plain run of the same command that exited 0
set -x
F="--steps 3 --warmup 3 --skip-build --skip-cpu --skip-small --skip-pipeline --e2e-messages 8"
python bench.py $F > gpurun_out/r02b_plain.json 2> gpurun_out/r02b_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed_apply/" --csv \
    --log-file gpurun_out/r02b_timed_apply_launches.csv python bench.py $F > gpurun_out/r02b_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
    --log-file gpurun_out/r02b_all_launches.csv python bench.py $F > gpurun_out/r02b_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:apply_window --nvtx --nvtx-include "timed_apply/" -c 1 \
    -o gpurun_out/r02b_apply_win -f python bench.py $F > gpurun_out/r02b_ncu3.log 2>&1
python tools/ncu_summary.py rep gpurun_out/r02b_apply_win.ncu-rep > gpurun_out/r02b_apply_win.ncu.txt
ncu -i gpurun_out/r02b_apply_win.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02b_apply_win.source.csv.gz
python tools/ncu_summary.py launches gpurun_out/r02b_timed_apply_launches.csv > gpurun_out/r02b_timed_apply_launches.txt
python tools/ncu_summary.py launches gpurun_out/r02b_all_launches.csv > gpurun_out/r02b_all_launches.txt
ls -la gpurun_out | tail -8
