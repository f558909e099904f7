# timing of the pipelined attention kernel under ncu (warm caches); SPLADE_B200_ATTN_DEBUG bits: 1 no ex2, 2 no bf16
# packs, 4 one P V MMA, 8 no S loads, 16 no V loads, 32 no Q/K loads, 64 no output stores (results WRONG, timing only)
export ONLY_ATTN=1
for d in ${ABLATIONS:-0}; do
  SPLADE_B200_ATTN_DEBUG=$d timeout 120 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -k regex:trunk_attn_tc2 -s 10 -c 6 --csv --log-file gpurun_out/abl_$d.csv python tools/time_trunk_kernels.py > /dev/null 2>&1
  echo "debug=$d: $(grep trunk_attn gpurun_out/abl_$d.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
done
