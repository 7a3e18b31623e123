# round 2: leak kernel variants with clock64 phase counters (tuning run)
tag=r2q
mkdir -p gpurun_out
export POLYCAP_B200_PHASES=1
for v in p1 p3 p4; do
  export POLYCAP_B200_LIB=$PWD/polycap_b200/variants/libpolycap_b200_$v.so
  for a in "30000 7" "2000 400"; do
    f=gpurun_out/${tag}_${v}_$(echo $a | tr ' ' '_').log
    timeout 300 python tools/leak_probe.py $a only > $f 2>&1
    echo "== $v $a"; grep -E "leak_calc|leak kernel|leak clocks|leak wave" $f | cut -c1-400
  done
done
