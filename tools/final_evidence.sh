#!/bin/bash
# Round-end evidence on one B200 (run under gpurun; every step under its own timeout, stdin closed):
#   tests, bench line, reference arm, ncu launch list, ncu --set full of the headline kernels and of the GEMMs, everyday-shape sweep
T=gpurun_out/final; mkdir -p $T
timeout 400 python -m pytest tests -m gpu -q < /dev/null > $T/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $T/pytest_gpu.log
timeout 300 python bench.py < /dev/null > $T/bench_n1.json 2> $T/bench_n1.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 < /dev/null > $T/bench_reference_arm.json 2>&1; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $T/bench_launches_ncu.csv python bench.py --steps 2 --warmup 1 < /dev/null > $T/ncu1.log 2>&1; echo "ncu launches rc=$?"
for k in reduce_row_kernel reduce_all_kernel; do
  timeout 300 ncu --set full --clock-control none -k regex:$k -s 2 -c 1 -o /tmp/$k -f python bench.py --steps 2 --warmup 1 --skip-others < /dev/null > $T/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
  ncu -i /tmp/$k.ncu-rep --page raw --csv > $T/ncu_headline_${k}_raw.csv 2>/dev/null
done
for c in c3d:dgemm_dmma c3s:gemm_tf32x3_kernel c4:gemm_tf32x3_fused2; do
  case=${c%%:*}; kern=${c##*:}
  timeout 300 bash tools/ncu_case.sh $case $kern 1 < /dev/null; cp gpurun_out/r2_ncu_${case}_raw.txt $T/ 2>/dev/null
done
timeout 500 python tools/cliff_hunt.py < /dev/null > $T/cliff_hunt.txt 2>&1; echo "cliff rc=$?"
timeout 100 python tools/call_overhead.py < /dev/null > $T/call_overhead.txt 2>&1
timeout 200 python tools/strip_perf.py < /dev/null > $T/strip_sweep.txt 2>&1
timeout 200 python tools/e2e_probe.py < /dev/null > $T/e2e_probe.json 2>$T/e2e_probe.err
echo done
