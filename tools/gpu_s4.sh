#!/bin/bash
# round-2 session 4 GPU steps (development aid): tools/gpu_s4.sh <tag> <step>...
tag=$1; shift
out=gpurun_out/$tag; mkdir -p $out
L=$PWD/seissol_b200/lib
QB="python tools/quick_bench.py --real 8 --cells 500000 --kernels ader,localwoader,neigh,all"
for s in "$@"; do
case $s in
pd2)
  for v in pd2 pd2n3; do
    SSB_LIB=$L/libseissol_b200_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $out/pytest_$v.log 2>&1; echo "exit $?" >> $out/pytest_$v.log
    SSB_LIB=$L/libseissol_b200_$v.so timeout 300 $QB > $out/qb_f64_$v.txt 2>&1
    SSB_ADER_GEN=4 SSB_LIB=$L/libseissol_b200_$v.so timeout 300 $QB > $out/qb_f64_${v}_ader4.txt 2>&1
  done
  timeout 300 $QB > $out/qb_f64_base.txt 2>&1 ;;
rsfast)
  timeout 600 python -m pytest tests/test_gpu_dr.py -m gpu -x -q -k "n_steps" > $out/pytest_dr.log 2>&1; echo "exit $?" >> $out/pytest_dr.log
  timeout 300 python bench.py --config dr_rsfvw_fast --steps 5 --no-cpu > $out/bench_dr_rsfvw_fast.json 2> $out/bench_dr_rsfvw_fast.err ;;
kt)
  for v in $KTV; do
    SSB_LIB=$L/libseissol_b200_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "float64" > $out/pytest_$v.log 2>&1; echo "exit $?" >> $out/pytest_$v.log
    SSB_LIB=$L/libseissol_b200_$v.so timeout 300 $QB > $out/qb_f64_$v.txt 2>&1
    SSB_ADER_GEN=4 SSB_LIB=$L/libseissol_b200_$v.so timeout 300 $QB > $out/qb_f64_${v}_ader4.txt 2>&1
  done ;;
final)
  timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $out/pytest.log 2>&1; echo "exit $?" >> $out/pytest.log
  timeout 600 python bench.py > $out/bench_elastic_f32.json 2> $out/bench_elastic_f32.err
  for c in elastic_f64 o4 visco dr_lsw dr_rsfvw dr_rsfvw_fast; do
    timeout 600 python bench.py --config $c --steps 10 > $out/bench_$c.json 2> $out/bench_$c.err
  done ;;
energy)
  timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_host_cpp.py -m gpu -x -q > $out/pytest_energy.log 2>&1; echo "exit $?" >> $out/pytest_energy.log
  timeout 300 python tools/quick_bench_stream.py > $out/qb_stream.txt 2>&1 ;;
energy2)
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "volume_energies" > $out/pytest_energy.log 2>&1; echo "exit $?" >> $out/pytest_energy.log
  timeout 300 python tools/quick_bench_stream.py > $out/qb_stream.txt 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu > $out/launches_bench.log 2>&1 ;;
ana)
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "analytical or nodal" > $out/pytest_ana.log 2>&1; echo "exit $?" >> $out/pytest_ana.log
  SSB_KGEN=3 SSB_KGEN5=1 timeout 300 $QB > $out/qb_f64_g3_local5.txt 2>&1
  SSB_KGEN=3 SSB_KGEN5=3 timeout 300 $QB > $out/qb_f64_g3_both5.txt 2>&1 ;;
lts2)
  (time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > $out/bench_lts2.json 2> $out/bench_lts2.err) 2> $out/bench_lts2.time
  (time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > $out/bench_ref2.json 2> $out/bench_ref2.err) 2> $out/bench_ref2.time
  timeout 600 python -m pytest tests/test_lts.py tests/test_host_cpp.py -m gpu -x -q > $out/pytest_lts2.log 2>&1; echo "exit $?" >> $out/pytest_lts2.log ;;
esac
done
for f in $out/*.log $out/*.txt $out/*.json $out/*.time; do echo "=== $f"; tail -c 1200 $f; done
