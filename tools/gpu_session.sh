#!/bin/bash
# One gpurun session: parity tests, micro-benchmarks, A/B kernel timings, bench.py, ncu captures.
# usage: tools/gpu_session.sh <tag> [steps...]   (steps: test ubench qb qbdr bench ncu)
tag=$1; shift
steps=${@:-"test qb bench ncu"}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/env.txt 2>&1
for s in $steps; do
case $s in
test)
  timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "exit $?" >> $out/pytest.log ;;
v4)
  NP7=$PWD/seissol_b200/lib/libseissol_b200_np7.so
  SSB_KGEN=4 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $out/pytest_v4.log 2>&1; echo "exit $?" >> $out/pytest_v4.log
  timeout 300 python tools/quick_bench.py > $out/qb_f32_v3.txt 2>&1
  SSB_KGEN=4 timeout 300 python tools/quick_bench.py > $out/qb_f32_v4.txt 2>&1
  [ -f $NP7 ] && SSB_LIB=$NP7 SSB_KGEN=4 timeout 300 python tools/quick_bench.py > $out/qb_f32_v4_np7.txt 2>&1
  timeout 300 python tools/quick_bench.py --real 8 --cells 500000 > $out/qb_f64_v4.txt 2>&1
  [ -f $NP7 ] && SSB_LIB=$NP7 timeout 300 python tools/quick_bench.py --real 8 --cells 500000 > $out/qb_f64_v4_np7.txt 2>&1
  SSB_KGEN=3 timeout 300 python tools/quick_bench.py --real 8 --cells 500000 > $out/qb_f64_v3.txt 2>&1 ;;
visco)
  timeout 900 python -m pytest tests/test_gpu_visco.py -m gpu -x -q > $out/pytest_visco.log 2>&1; echo "exit $?" >> $out/pytest_visco.log
  timeout 300 python tools/quick_bench_visco.py > $out/qb_visco.txt 2>&1
  timeout 300 python tools/quick_bench_visco.py --real 8 --cells 500000 > $out/qb_visco_f64.txt 2>&1 ;;
lts)
  timeout 900 python -m pytest tests/test_lts.py -m gpu -x -q > $out/pytest_lts.log 2>&1; echo "exit $?" >> $out/pytest_lts.log
  timeout 600 python tools/quick_bench_lts.py > $out/qb_lts.txt 2>&1 ;;
part2lts)
  timeout 600 python -m pytest tests/test_lts.py -m gpu -x -q > $out/pytest_lts2.log 2>&1; echo "exit $?" >> $out/pytest_lts2.log
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 --lts 3 > $out/bench_lts2.json 2> $out/bench_lts2.err
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_part2.json 2> $out/bench_part2.err ;;
ab4)
  for v in c4 c3; do
    L=$PWD/seissol_b200/lib/libseissol_b200_$v.so
    [ -f $L ] && SSB_LIB=$L timeout 300 python tools/quick_bench.py > $out/qb_$v.txt 2>&1
  done
  timeout 300 python tools/quick_bench.py > $out/qb_base.txt 2>&1 ;;
ncu3)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_visco' --launch-skip 6 -c 3 -f -o $out/full_visco \
      python tools/quick_bench_visco.py --cells 200000 --steps 1 --kernels all > $out/ncu_visco.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_local|k_neigh' --launch-skip 6 -c 3 -f -o $out/full_f64 \
      python tools/quick_bench.py --real 8 --cells 148000 --steps 1 --kernels all > $out/ncu_f64.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_dr_interp|k_friction|k_neigh_dr' --launch-skip 5 -c 4 -f -o $out/full_dr \
      python tools/quick_bench_dr.py --steps 1 > $out/ncu_dr.log 2>&1 ;;
occ)
  timeout 300 python tools/quick_bench.py > $out/qb_occ2.txt 2>&1
  SSB_SMEM_PAD=20000 timeout 300 python tools/quick_bench.py > $out/qb_occ1.txt 2>&1
  SSB_SMEM_PAD=20000 SSB_PERSISTENT=1 timeout 300 python tools/quick_bench.py > $out/qb_occ1_persist.txt 2>&1 ;;
qbf64)
  timeout 300 python tools/quick_bench.py --real 8 --cells 500000 > $out/qb_f64.txt 2>&1 ;;
part)
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_part2.json 2> $out/bench_part2.err
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-overlap > $out/bench_part2_noov.json 2> $out/bench_part2_noov.err ;;
ls)
  timeout 600 tools/ubench/ifetch4 > $out/ifetch4.txt 2>&1
  timeout 300 python tools/quick_bench.py > $out/qb_base.txt 2>&1
  for v in $LSVARIANTS; do
    L=$PWD/seissol_b200/lib/libseissol_b200_$v.so
    [ -f $L ] && SSB_LIB=$L timeout 300 python tools/quick_bench.py > $out/qb_$v.txt 2>&1
  done
  for v in $LSTEST; do
    L=$PWD/seissol_b200/lib/libseissol_b200_$v.so
    [ -f $L ] && SSB_LIB=$L timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $out/pytest_$v.log 2>&1; echo "exit $?" >> $out/pytest_$v.log
  done ;;
fp)
  timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $out/pytest_parity.log 2>&1; echo "exit $?" >> $out/pytest_parity.log
  timeout 300 python tools/quick_bench.py > $out/qb_base.txt 2>&1
  SSB_FACEPROJ=0 timeout 300 python tools/quick_bench.py > $out/qb_nofp.txt 2>&1
  for v in $LSVARIANTS; do
    L=$PWD/seissol_b200/lib/libseissol_b200_$v.so
    [ -f $L ] && SSB_LIB=$L timeout 300 python tools/quick_bench.py > $out/qb_$v.txt 2>&1
  done ;;
tma)
  timeout 300 tools/ubench/tma_rate > $out/tma_rate.txt 2>&1 ;;
ncufp)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_local|k_neigh' --launch-skip 6 -c 3 -f -o $out/full_fp \
      python tools/quick_bench.py --cells 296000 --steps 1 --kernels all > $out/ncu_full_fp.log 2>&1 ;;
g5)
  timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $out/pytest_parity.log 2>&1; echo "exit $?" >> $out/pytest_parity.log
  timeout 300 python tools/quick_bench.py > $out/qb_base.txt 2>&1
  SSB_KGEN5=0 timeout 300 python tools/quick_bench.py > $out/qb_gen3.txt 2>&1
  for v in $LSVARIANTS; do
    L=$PWD/seissol_b200/lib/libseissol_b200_$v.so
    [ -f $L ] && SSB_LIB=$L timeout 300 python tools/quick_bench.py > $out/qb_$v.txt 2>&1
  done ;;
qbv)
  timeout 300 python tools/quick_bench.py --kernels ader,localwoader,neigh,all > $out/qb_base.txt 2>&1
  for v in $LSVARIANTS; do
    L=$PWD/seissol_b200/lib/libseissol_b200_$v.so
    [ -f $L ] && SSB_LIB=$L timeout 300 python tools/quick_bench.py --kernels ader,localwoader,neigh,all > $out/qb_$v.txt 2>&1
  done ;;
fp2)
  timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $out/pytest_parity.log 2>&1; echo "exit $?" >> $out/pytest_parity.log
  timeout 300 python tools/quick_bench.py --kernels ader,localwoader,neigh,all > $out/qb_base.txt 2>&1
  SSB_NEIGHFP=1 timeout 300 python tools/quick_bench.py --kernels neigh,all > $out/qb_fp1.txt 2>&1
  timeout 300 python tools/quick_bench.py --kernels neigh,all --nosort > $out/qb_nosort.txt 2>&1
  timeout 300 python tools/quick_bench.py --real 8 --cells 500000 --kernels ader,localwoader,neigh,all > $out/qb_f64.txt 2>&1
  SSB_FACEPROJ=0 timeout 300 python tools/quick_bench.py --real 8 --cells 500000 --kernels ader,localwoader,neigh,all > $out/qb_f64_nofp.txt 2>&1
  for v in $LSVARIANTS; do
    L=$PWD/seissol_b200/lib/libseissol_b200_$v.so
    [ -f $L ] && SSB_LIB=$L timeout 300 python tools/quick_bench.py --kernels neigh,all > $out/qb_$v.txt 2>&1
  done ;;
pk2)
  for v in $LSVARIANTS; do
    L=$PWD/seissol_b200/lib/libseissol_b200_$v.so
    [ -f $L ] && SSB_FACEPROJ=0 SSB_LIB=$L timeout 300 python tools/quick_bench.py --kernels ader,localwoader,neigh,all > $out/qb_$v.txt 2>&1
  done ;;
f64v)
  timeout 300 tools/ubench/tma_rate > $out/tma_rate.txt 2>&1
  SSB_KGEN5=1 timeout 300 python tools/quick_bench.py --real 8 --cells 500000 --kernels ader,localwoader,neigh,all > $out/qb_f64_g5local.txt 2>&1
  SSB_KGEN5=3 timeout 300 python tools/quick_bench.py --real 8 --cells 500000 --kernels ader,localwoader,neigh,all > $out/qb_f64_g5both.txt 2>&1
  SSB_KGEN=3 timeout 300 python tools/quick_bench.py --real 8 --cells 500000 --kernels ader,localwoader,neigh,all > $out/qb_f64_g3.txt 2>&1
  timeout 300 python tools/quick_bench.py --order 4 --kernels ader,localwoader,neigh,all > $out/qb_o4.txt 2>&1
  SSB_FACEPROJ=0 SSB_KGEN5=0 timeout 300 python tools/quick_bench.py --order 4 --kernels ader,localwoader,neigh,all > $out/qb_o4_r1.txt 2>&1
  timeout 300 python tools/quick_bench.py --order 5 --kernels ader,localwoader,neigh,all > $out/qb_o5.txt 2>&1
  SSB_FACEPROJ=0 SSB_KGEN5=0 timeout 300 python tools/quick_bench.py --order 5 --kernels ader,localwoader,neigh,all > $out/qb_o5_r1.txt 2>&1 ;;
ncuf64)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_local' --launch-skip 4 -c 2 -f -o $out/full_f64fp \
      python tools/quick_bench.py --real 8 --cells 148000 --steps 1 --kernels all > $out/ncu_f64fp.log 2>&1
  SSB_FACEPROJ=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_local' --launch-skip 4 -c 2 -f -o $out/full_f64nofp \
      python tools/quick_bench.py --real 8 --cells 148000 --steps 1 --kernels all > $out/ncu_f64nofp.log 2>&1 ;;
benchall)
  timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "exit $?" >> $out/pytest.log
  timeout 600 python bench.py > $out/bench_elastic_f32.json 2> $out/bench_elastic_f32.err
  for c in elastic_f64 o4 visco dr_lsw dr_rsfvw; do
    timeout 600 python bench.py --config $c --steps 10 > $out/bench_$c.json 2> $out/bench_$c.err
  done
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err ;;
bench2)
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 > $out/bench_lts2.json 2> $out/bench_lts2.err
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 10 --warmup 3 --lts 1 > $out/bench_gts2.json 2> $out/bench_gts2.err
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > $out/bench_ref2.json 2> $out/bench_ref2.err
  timeout 900 python -m pytest tests/test_lts.py -m gpu -x -q > $out/pytest_lts.log 2>&1; echo "exit $?" >> $out/pytest_lts.log ;;
drtests)
  timeout 1500 python -m pytest tests/test_gpu_dr.py tests/test_lts.py tests/test_gpu_parity.py tests/test_host_cpp.py -m gpu -x -q > $out/pytest_dr.log 2>&1; echo "exit $?" >> $out/pytest_dr.log
  timeout 600 python bench.py --config dr_rsfvw --steps 10 > $out/bench_dr_rsfvw.json 2> $out/bench_dr_rsfvw.err
  timeout 300 python tools/quick_bench_lts.py > $out/qb_lts.txt 2>&1
  SSB_FACEPROJ=0 timeout 300 python tools/quick_bench_lts.py > $out/qb_lts_nofp.txt 2>&1 ;;
tc)
  SSB_ADER_TC=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ader" > $out/pytest_tc.log 2>&1; echo "exit $?" >> $out/pytest_tc.log
  SSB_ADER_TC=1 timeout 300 python tools/quick_bench.py --kernels ader,all > $out/qb_tc.txt 2>&1
  timeout 300 python tools/quick_bench.py --kernels ader,all > $out/qb_simt.txt 2>&1
  SSB_ADER_TC=1 timeout 300 python tools/quick_bench.py --order 4 --kernels ader,all > $out/qb_tc_o4.txt 2>&1 ;;
ncutc)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_ader_tc' --launch-skip 2 -c 1 -f -o $out/full_adertc \
      env SSB_ADER_TC=1 python tools/quick_bench.py --cells 296000 --steps 1 --kernels ader > $out/ncu_adertc.log 2>&1 ;;
drtc2)
  SSB_DR_TC=1 timeout 600 python -m pytest tests/test_gpu_dr.py -m gpu -q -k "godunov or chain or n_steps" > $out/pytest_drtc.log 2>&1; echo "exit $?" >> $out/pytest_drtc.log
  SSB_DR_TC=1 timeout 300 python tools/quick_bench_dr.py > $out/qbdr_tc.txt 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_dr_interp_tc' --launch-skip 2 -c 1 -f -o $out/full_drtc \
      env SSB_DR_TC=1 python tools/quick_bench_dr.py --steps 1 > $out/ncu_drtc.log 2>&1 ;;
nfvar)
  for v in nf14 nf16 nf10; do
    L=$PWD/seissol_b200/lib/libseissol_b200_$v.so
    [ -f $L ] && SSB_LIB=$L timeout 300 python tools/quick_bench.py --kernels neigh,all > $out/qb_$v.txt 2>&1
  done
  timeout 300 python tools/quick_bench.py --kernels neigh,all > $out/qb_base.txt 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_local5|k_neigh_fp' --launch-skip 2 -c 2 -f -o $out/full_l5 \
      python tools/quick_bench.py --cells 296000 --steps 1 --kernels all > $out/ncu_l5.log 2>&1 ;;
umma)
  timeout 60 tools/ubench/umma_test > $out/umma_test.txt 2>&1; echo "exit $?" >> $out/umma_test.txt
  timeout 120 tools/ubench/fma_peak > $out/fma_peak.txt 2>&1 ;;
drtc)
  timeout 60 tools/ubench/umma_test > $out/umma_test.txt 2>&1; echo "exit $?" >> $out/umma_test.txt
  timeout 120 tools/ubench/fma_peak > $out/fma_peak.txt 2>&1
  timeout 900 python -m pytest tests/test_gpu_dr.py -m gpu -x -q -k "godunov or chain or n_steps" > $out/pytest_drtc.log 2>&1; echo "exit $?" >> $out/pytest_drtc.log
  SSB_DR_TC=1 timeout 300 python tools/quick_bench_dr.py > $out/qbdr_tc.txt 2>&1
  timeout 300 python tools/quick_bench_dr.py > $out/qbdr_simt.txt 2>&1
  timeout 900 python -m pytest tests/test_lts.py tests/test_gpu_parity.py tests/test_host_cpp.py tests/test_gpu_visco.py tests/test_planewave.py -m gpu -x -q > $out/pytest_rest.log 2>&1; echo "exit $?" >> $out/pytest_rest.log ;;
ubench)
  timeout 300 tools/ubench/ifetch3 > $out/ifetch3.txt 2>&1 ;;
qb)
  timeout 300 python tools/quick_bench.py > $out/qb.txt 2>&1
  SSB_PERSISTENT=2 timeout 300 python tools/quick_bench.py > $out/qb_persist2.txt 2>&1
  timeout 300 python tools/quick_bench.py --real 8 --cells 500000 > $out/qb_f64.txt 2>&1 ;;
qbdr)
  timeout 300 python tools/quick_bench_dr.py > $out/qbdr.txt 2>&1
  timeout 300 python tools/quick_bench_dr.py --real 8 > $out/qbdr_f64.txt 2>&1 ;;
bench)
  timeout 900 python bench.py > $out/bench.json 2> $out/bench.err ;;
benchref)
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err ;;
ncu)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu > $out/launches_bench.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_local|k_neigh' --launch-skip 9 -c 3 -f -o $out/full \
      python tools/quick_bench.py --cells 296000 --steps 1 --kernels all > $out/ncu_full.log 2>&1 ;;
ncu2)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_local|k_neigh' --launch-skip 6 -c 3 -f -o $out/full_f32 \
      python tools/quick_bench.py --cells 296000 --steps 1 --kernels all > $out/ncu_full_f32.log 2>&1
  SSB_KGEN=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_local|k_neigh' --launch-skip 6 -c 3 -f -o $out/full_f64v3 \
      python tools/quick_bench.py --real 8 --cells 148000 --steps 1 --kernels all > $out/ncu_full_f64v3.log 2>&1 ;;
ncudr)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_dr_interp|k_friction|k_neigh_dr' --launch-skip 5 -c 4 -f -o $out/full_dr \
      python tools/quick_bench_dr.py --steps 1 > $out/ncu_full_dr.log 2>&1 ;;
esac
done
# .ncu-rep files above 30 MB would push gpurun_out over its 64 MiB limit: keep the CSV pages instead
for rep in $out/*.ncu-rep; do
  [ -f "$rep" ] || continue
  ncu -i $rep --page raw --csv > ${rep%.ncu-rep}_raw.csv 2>/dev/null
  if [ $(stat -c %s $rep) -gt 30000000 ]; then
    ncu -i $rep --page source --csv > ${rep%.ncu-rep}_source.csv 2>/dev/null
    gzip -f ${rep%.ncu-rep}_source.csv
    rm -f $rep
  fi
done
ls -la $out
for f in $out/*.log $out/*.txt $out/*.json $out/*.err; do echo "=== $f"; tail -c 1500 $f; done
