# round 2: phased leak kernel, block size / blocks per SM
tag=r2x
mkdir -p gpurun_out
export POLYCAP_B200_PHASES=1
for cfg in "ph2 9" "ph 1" "ph2 2"; do
  set -- $cfg
  f=gpurun_out/${tag}_leak_$1_$2_30000_7.log
  POLYCAP_B200_LEAK_BLOCKS_PER_SM=$2 POLYCAP_B200_LIB=$PWD/polycap_b200/variants/libpolycap_b200_$1.so timeout 90 python tools/leak_probe.py 30000 7 only > $f 2>&1
  echo "== leak variant $cfg"; grep -E "leak_calc|leak kernel|leak clocks|leak wave" $f | cut -c1-330
done
f=gpurun_out/${tag}_leak_ph_2_20000_400.log
POLYCAP_B200_LIB=$PWD/polycap_b200/variants/libpolycap_b200_ph.so timeout 120 python tools/leak_probe.py 20000 400 only > $f 2>&1
echo "== leak variant ph, 20000 400"; grep -E "leak_calc|leak kernel|leak clocks" $f | cut -c1-330
