#!/bin/bash
# Round-2 evidence run on ONE B200 (through gpurun): bench line, reference arm, launch list, ncu --set full captures, config 2 at length.
# Every ncu command runs only after the same command has exited 0 without the tool.
# (compute-sanitizer is closed on this pool — see profiles/r02_sanitizer.txt; tools/sanitize_smoke.py is the small-shape program it would run.)
set -x
O=gpurun_out
python -m pytest tests -x -q -m gpu > $O/r02_gpu_tests.txt 2>&1; tail -2 $O/r02_gpu_tests.txt; python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.txt 2>&1; cat $O/r02_smoke.txt
python bench.py > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err || exit 1
python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_reference_n1.json 2> $O/r02_bench_reference_n1.err
python tools/profile_single_frame.py 3 50 > $O/r02_profile_single_frame.txt 2>&1 || exit 1
python tools/sanitize_smoke.py > $O/r02_sanitize_plain.txt 2>&1 || exit 1
# launch list of one bench window (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_bench_launches_ncu.csv -k regex:"^k_|liorf" -c 2000 python bench.py --steps 5 --warmup 3 --no-extras > $O/ncu_launch.log 2>&1
# ncu --set full: solver and the one-kernel VoxelGrid on kitti64_single
ncu --set full --clock-control none --import-source on -k regex:k_scan2map_persistent -s 3 -c 1 -o $O/r02_scan2map python tools/profile_single_frame.py 3 50 > $O/ncu_s2m.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_vg_fused -s 52 -c 1 -o $O/r02_vg_fused python tools/profile_single_frame.py 3 50 > $O/ncu_vgf.log 2>&1
# config 2 at length
python tools/run_kitti05_full.py 2761 300 $O/r02_kitti05_full.json > $O/r02_kitti05_full.log 2>&1
ls -la $O | tail -20
