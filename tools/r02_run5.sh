#!/bin/bash
# round 2, GPU call 5: ncu of the cfg5 / cfg4 kernels (DRAM traffic, own counters), contact-kernel profile after the rolled SAT,
# where the contact kernel's time goes, repeated runs of the test that failed once
O=gpurun_out/r02_e
mkdir -p $O
for i in 1 2 3 4; do timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "contacts or exhaustive or edge or host_api or both_objects" > $O/pytest_rep$i.log 2>&1; echo "rep $i rc=$?"; done
timeout 300 python tools/split_timing.py 2>&1 | tee $O/split_timing.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"collide_ordered_kernel" -c 1 -f -o $O/full_contacts python tools/profile_run.py --workload contacts --poses 1000000 --traversal 3 --launches 1 > $O/full_contacts.log 2>&1
python tools/ncu_summary.py $O/full_contacts.ncu-rep > $O/full_contacts.summary.txt 2>&1
python tools/ncu_by_function.py $O/full_contacts.ncu-rep >> $O/full_contacts.summary.txt 2>&1
head -45 $O/full_contacts.summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"collide_front_kernel|distance_warp_kernel" -c 3 -f -o $O/full_cfg5 python tools/profile_big.py --workload cfg5 --poses 100000 > $O/full_cfg5.log 2>&1
tail -4 $O/full_cfg5.log
python tools/ncu_summary.py $O/full_cfg5.ncu-rep > $O/full_cfg5.summary.txt 2>&1
grep -E "^==|time_duration|dram__bytes|lts__t_sector_hit|l1tex__t_sector_hit|issue_active|warps_active|stall|inst_executed.sum|registers" $O/full_cfg5.summary.txt
timeout 600 ncu --set full --clock-control none -k regex:"collide_front_kernel" -c 1 -f -o $O/full_cfg4 python tools/profile_big.py --workload cfg4 --poses 1000000 > $O/full_cfg4.log 2>&1
tail -3 $O/full_cfg4.log
python tools/ncu_summary.py $O/full_cfg4.ncu-rep > $O/full_cfg4.summary.txt 2>&1
grep -E "^==|time_duration|dram__bytes|lts__t_sector_hit|l1tex__t_sector_hit|issue_active|warps_active|stall|inst_executed.sum" $O/full_cfg4.summary.txt
