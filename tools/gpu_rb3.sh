cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/profile_refiner.py --sweep 1 > gpurun_out/rb_sweep.log 2>&1
for cfg in "rb_dw p1_s1 dw_s1" "rb_pw p1_s1 pw_s1" "rb_dw p1_s16 dw_s16" "rb_pw p1_s16 pw_s16" "rb_dw p1_s2 dw_s2"; do
  set -- $cfg
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$1 --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2_full_$3 python tools/profile_refiner.py --shape $2 --b 64 > gpurun_out/ncu_$3.log 2>&1
done
cat gpurun_out/rb_sweep.log; ls -la gpurun_out/*.ncu-rep | tail -8
