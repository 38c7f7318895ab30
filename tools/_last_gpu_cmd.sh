python tools/check_resident.py 61960 4099 2>&1 | grep -E "rows|fp64|phases" > gpurun_out/r2u_check.log
KIND=c4 python tools/check_resident.py 4099 2>&1 | grep -E "rows|fp64|phases" >> gpurun_out/r2u_check.log
KIND=go9 python tools/check_resident.py 2033 2>&1 | grep -E "rows|fp64|phases" >> gpurun_out/r2u_check.log
cat gpurun_out/r2u_check.log
for v in fused_heads new nostem9 fused_heads new; do
  unset SPRL_B200_LIB SPRL_EVALNET_NO_STEM9
  if [ "$v" = fused_heads ]; then export SPRL_B200_LIB=sprl_b200/lib/variants/fused_heads.so; fi
  if [ "$v" = nostem9 ]; then export SPRL_EVALNET_NO_STEM9=1; fi
  echo "== $v"; REPS=400 python tools/check_resident.py 61960 2>&1 | grep -E "rows|resident: sm clock"
done > gpurun_out/r2u_ab.log 2>&1
unset SPRL_B200_LIB SPRL_EVALNET_NO_STEM9
cat gpurun_out/r2u_ab.log
SPRL_B200_LIB=sprl_b200/lib/variants/trace.so SPRL_EVALNET_TRACE=1 python tools/check_resident.py 61960 > gpurun_out/r2u_trace.log 2>&1
tools/search_variants.sh base w2b16 w2b20 w2b24 w1b24 w1b16 base > gpurun_out/r2u_search_variants.log 2>&1
cat gpurun_out/r2u_search_variants.log
