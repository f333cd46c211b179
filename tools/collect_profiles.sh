#!/bin/bash
# Round-2 evidence run on ONE B200 (through gpurun): bench line, launch list, ncu --set full captures, sanitizer logs, config 2 at length.
# Every ncu / sanitizer command runs only after the same command has exited 0 without the tool.
set -x
O=gpurun_out
python bench.py > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err || exit 1
python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_reference_n1.json 2> $O/r02_bench_reference_n1.err
python tools/profile_single_frame.py 3 50 > $O/r02_profile_single_frame.txt 2>&1 || exit 1
python tools/profile_front.py 3 > $O/r02_profile_front.txt 2>&1 || exit 1
python tools/profile_sc.py 100000 32768 3 > $O/r02_profile_sc.txt 2>&1 || exit 1
# launch list of one bench window (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_bench_launches_ncu.csv -k regex:"^k_|liorf" -c 2000 python bench.py --steps 5 --warmup 3 --no-extras > $O/ncu_launch.log 2>&1
# ncu --set full: solver on kitti64_single, the streaming kernels, the two ScanContext kernels
ncu --set full --clock-control none --import-source on -k regex:k_scan2map_persistent -s 3 -c 1 -o $O/r02_scan2map python tools/profile_single_frame.py 3 50 > $O/ncu_s2m.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_deskew_points|k_first_kept" -s 2 -c 2 -o $O/r02_deskew python tools/profile_front.py 2 > $O/ncu_dk.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_radix_pass|k_vg_centroid|k_vg_keys|k_vg_minmax|k_radix_hist" -s 470 -c 16 -o $O/r02_voxelgrid python tools/profile_front.py 2 > $O/ncu_vg.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_transform_concat|k_grid_scatter|k_grid_count" -s 3 -c 3 -o $O/r02_map python tools/profile_front.py 2 > $O/ncu_map.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sc_tensor -s 2 -c 1 -o $O/r02_sc_tensor python tools/profile_sc.py 100000 32768 2 > $O/ncu_sct.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sc_distance_bulk -s 2 -c 1 -o $O/r02_sc_distance_bulk python tools/profile_sc.py 100000 32768 2 > $O/ncu_scd.log 2>&1
# sanitizer passes over the small-shape smoke
python tools/sanitize_smoke.py > $O/r02_sanitize_plain.txt 2>&1 || exit 1
compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py > $O/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" >> $O/r02_sanitizer_memcheck.txt
compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_smoke.py > $O/r02_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" >> $O/r02_sanitizer_racecheck.txt
# config 2 at length
python tools/run_kitti05_full.py 2761 300 $O/r02_kitti05_full.json > $O/r02_kitti05_full.log 2>&1
ls -la $O | tail -40
