set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/t.log
export DC_CUDA_GRAPH=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 324 -c 162 --csv --log-file gpurun_out/launches_r1_final.csv python tools/profile_forward.py --warm 2 --iters 1 > gpurun_out/ncu_l.log 2>&1
for spec in res2a_2b:3 res4b7_2b:59 res4b7_2c:60 res5b_2b:150; do
  name=${spec%%:*}; idx=${spec##*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_igemm --launch-skip $((314 + idx)) --launch-count 1 -f -o gpurun_out/prof_final_$name python tools/profile_forward.py --warm 2 --iters 1 > gpurun_out/ncu_$name.log 2>&1
done
cat gpurun_out/t.log; ls -la gpurun_out/prof_final_*; tail -3 gpurun_out/ncu_l.log
