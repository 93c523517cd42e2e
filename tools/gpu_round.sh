#!/bin/bash
# One GPU-box visit: parity tests, the bench line (both arms), an ncu launch list and full captures of the
# two dominant kernels.  Everything lands in gpurun_out/<tag>_*.   usage: tools/gpu_round.sh <tag> [what...]
set -u
TAG=${1:-r01}; shift || true
WHAT=${*:-tests bench ref launches ncu}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
for w in $WHAT; do
  case $w in
    tests)
      timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
      echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log; tail -3 $OUT/${TAG}_pytest_gpu.log ;;
    debugtc)
      DPDF_OPTIONS=intra_tc=1 timeout 600 python tools/gpu_debug.py dpdfnet2 > $OUT/${TAG}_debug_tc.log 2>&1; tail -25 $OUT/${TAG}_debug_tc.log ;;
    quick)   # kernel-time table at two batch sizes, no ladder
      for b in 1024 8192; do
        timeout 600 python bench.py --steps 50 --warmup 10 --no-ladder --no-extras --latency-hops 50 --batch $b --cpu-hops 2 --cpu-batch 16 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('B=$b', round(d['ms_per_step'], 4), 'ms/hop', int(d['value']), 'sf/s', d['kernel_ms'])
    else:
        print(l, end='')
" | tee -a $OUT/${TAG}_quick.log
      done ;;
    lanes)   # hop time against the number of lanes
      for b in 1024 2048 8192; do for L in 1 2 4 8; do
        timeout 600 python bench.py --steps 60 --warmup 10 --no-ladder --batch $b --lanes $L --profile-only 2>&1 | sed "s/^/B=$b lanes=$L /" | tee -a $OUT/${TAG}_lanes.log
      done; done ;;
    configs)   # BASELINE.json configs[2..4]: other model sizes / 48 kHz (device-resident throughput + kernel table)
      for cfg in "dpdfnet8 4096" "dpdfnet2_48khz_hr 2048" "dpdfnet8_48khz_hr 2048" "dpdfnet2 1024"; do
        set -- $cfg
        timeout 600 python bench.py --steps 40 --warmup 10 --no-ladder --no-extras --latency-hops 50 --model $1 --batch $2 --cpu-hops 2 --cpu-batch 16 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$1 B=$2', round(d['ms_per_step'], 4), 'ms/hop', int(d['value']), 'sf/s', 'realtime', int(d['realtime_streams']), d['kernel_ms'])
    else:
        print(l, end='')
" | tee -a $OUT/${TAG}_configs.log
      done ;;
    resample)  # device resampler throughput (HBM-bound: 4 B in + 4*up/down B out per input sample)
      timeout 300 python - <<'PYEOF' 2>&1 | tee $OUT/${TAG}_resample.log
import torch, time
from dpdfnet_b200.resample import BatchResampler
for sr_in, sr_out in ((48000, 16000), (16000, 48000), (44100, 16000)):
    B, n = 2048, 48000
    rs = BatchResampler(sr_in, sr_out, B)
    x = torch.randn(B, n, device="cuda")
    for _ in range(3): rs.resample(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): y = rs.resample(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = (x.numel() + y.numel()) * 4 / 1e9
    print(f"{sr_in}->{sr_out}: {B} streams x {n} samples in {ms:.3f} ms = {B * n / ms / 1e6:.1f} G input samples/s, {gb / ms * 1e3:.0f} GB/s of algorithmic traffic")
PYEOF
      ;;
    smoke)
      timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log ;;
    bench)
      timeout 900 python bench.py ${BENCH_ARGS:-} > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
      echo "bench exit $?"; tail -c 3000 $OUT/${TAG}_bench.json ;;
    ref)
      timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
      echo "ref exit $?"; cat $OUT/${TAG}_bench_ref.json ;;
    launches)
      # 5 warm-up hops skipped by -s (launches per hop printed by the bench), one hop listed; graph off so
      # every node is a plain launch
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 120 --csv \
        --log-file $OUT/${TAG}_launches.csv python bench.py --steps 8 --warmup 8 --no-graph --profile-only --lanes 1 ${BENCH_ARGS:-} \
        > $OUT/${TAG}_launches.log 2>&1
      echo "launches exit $?" ;;
    ncu)
      for k in ${NCU_KERNELS:-k_dprnn_intra k_dprnn_post_tc}; do
        timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 2 \
          -f -o $OUT/${TAG}_$k python bench.py --steps 4 --warmup 4 --no-graph --profile-only --lanes 1 ${BENCH_ARGS:-} \
          > $OUT/${TAG}_ncu_$k.log 2>&1
        echo "ncu $k exit $?"
      done ;;
    ncubig)   # throughput regime: full captures of the dominant kernels at NCU_BATCH streams (one chain) for NCU_MODEL
      for k in ${NCU_KERNELS:-k_dprnn_post_tc k_sepconv_tc k_dprnn_intra_tc}; do
        timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 2 \
          -f -o $OUT/${TAG}_${NCU_MODEL:-dpdfnet4}_B${NCU_BATCH:-16384}_$k python bench.py --steps 3 --warmup 3 --no-graph --profile-only --lanes 1 \
          --model ${NCU_MODEL:-dpdfnet4} --batch ${NCU_BATCH:-16384} > $OUT/${TAG}_ncubig_$k.log 2>&1
        echo "ncubig $k exit $?"
      done ;;
    ncuall)
      timeout 1500 ncu --set full --clock-control none --import-source on -s 300 -c 60 \
        -f -o $OUT/${TAG}_allkernels python bench.py --steps 8 --warmup 8 --no-graph --profile-only --lanes 1 \
        > $OUT/${TAG}_ncu_all.log 2>&1
      echo "ncuall exit $?" ;;
  esac
done
ls -la $OUT | tail -20
