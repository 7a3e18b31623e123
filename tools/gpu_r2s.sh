# round 2: leak kernel, resident warps per SM (instruction-fetch-bound: does throughput depend on them at all?)
tag=r2s
mkdir -p gpurun_out
export POLYCAP_B200_PHASES=1
for cfg in "t32 1" "t32 2" "t32 4" "p1 1" "p1 2"; do
  set -- $cfg
  export POLYCAP_B200_LIB=$PWD/polycap_b200/variants/libpolycap_b200_$1.so POLYCAP_B200_LEAK_BLOCKS_PER_SM=$2
  for a in "30000 7" "2000 400"; do
    f=gpurun_out/${tag}_$1_$2_$(echo $a | tr ' ' '_').log
    timeout 300 python tools/leak_probe.py $a only > $f 2>&1
    echo "== $cfg : $a"; grep -E "leak_calc|leak kernel|leak clocks" $f | cut -c1-330
  done
done
