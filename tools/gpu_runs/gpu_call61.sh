# ncu capture of k_round after the instruction diet (same command as r1q)
python tools/profile_search.py 16384 6 ext 1 1500 > gpurun_out/plain_prof_r1v.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_round -s 1502 -c 2 -o gpurun_out/prof_round_r1v -f python tools/profile_search.py 16384 6 ext 1 1500 > gpurun_out/ncu_prof_r1v.log 2>&1; echo rc=$?
tail -1 gpurun_out/plain_prof_r1v.log
