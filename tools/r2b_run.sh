#!/bin/bash
# second evidence run of round 2: sharded stage 2 under ncu (ranks emulated on one GPU), VoxelGrid kernels under ncu
set -x
O=gpurun_out
python tools/profile_sc_shard.py 100000 32768 8 6 10 > $O/r02_profile_sc_shard_q32k.txt 2>&1 || exit 1
ncu --section SpeedOfLight --section WarpStateStats --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --section SchedulerStats --section SourceCounters --import-source on --clock-control none -k regex:k_sc_distance_bulk -c 3 -o $O/r02_sc_distance_bulk_sharded python tools/profile_sc_shard.py 100000 32768 8 1 1 > $O/ncu_scd_sh.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_front_launches_ncu.csv python tools/profile_front.py 2 > $O/ncu_front_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_radix_pass|k_vg_centroid|k_vg_keys|k_vg_minmax|k_radix_hist" -s 16 -c 8 -o $O/r02_voxelgrid python tools/profile_front.py 2 > $O/ncu_vg.log 2>&1
tail -3 $O/ncu_vg.log
cat $O/r02_profile_sc_shard_q32k.txt | tail -12
