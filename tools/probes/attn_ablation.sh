export ONLY_ATTN=1
for cfg in "16384 127" "8192 127" "4096 127" "8192 0" "4096 0"; do
  set -- $cfg
  T=$1 SPLADE_B200_ATTN_DEBUG=$2 timeout 120 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -k regex:trunk_attn_tc2 -s 10 -c 6 --csv --log-file gpurun_out/abl_$1_$2.csv python tools/time_trunk_kernels.py > /dev/null 2>&1
  echo "T=$1 debug=$2: $(grep trunk_attn gpurun_out/abl_$1_$2.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
done
