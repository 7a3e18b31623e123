# round 2: ncu --set full capture of the leak kernel (small run: 3000 exit photons, 40 energies)
tag=r2w
mkdir -p gpurun_out
timeout 120 python tools/leak_probe.py 3000 40 only > gpurun_out/${tag}_plain.log 2>&1 || exit 1
tail -1 gpurun_out/${tag}_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pcb_leak_kernel -c 1 -f -o gpurun_out/prof_${tag}_leak python tools/leak_probe.py 3000 40 only > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log; ls -la gpurun_out/prof_${tag}_leak.ncu-rep
