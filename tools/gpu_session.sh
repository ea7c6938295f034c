#!/bin/bash
# One gpurun session: tests + bench + profiles; every part has its own timeout and log under gpurun_out/.
# usage: tools/gpu_session.sh <tag> [parts...]   parts: selftest tests bench ncu_wgrad launches
TAG=${1:-r02}; shift
PARTS=${@:-"selftest tests bench"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
for part in $PARTS; do
  case $part in
    selftest)
      timeout 300 tests/native/gemm_selftest perf > gpurun_out/${TAG}_selftest_perf.log 2>&1
      echo "[selftest perf] rc=$?"; grep PERF gpurun_out/${TAG}_selftest_perf.log | tail -20 ;;
    tests)
      CRIS_B200_EXPERIMENTAL=1 timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -rA -s \
        > gpurun_out/${TAG}_tests.log 2>&1
      echo "[tests] rc=$?"; grep -E "^(PASSED|FAILED|ERROR|SKIPPED)|passed|failed" gpurun_out/${TAG}_tests.log | tail -90 ;;
    tests_ops)
      CRIS_B200_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_parity_gpu.py -q --no-header -p no:cacheprovider -rA \
        > gpurun_out/${TAG}_tests_ops.log 2>&1
      echo "[tests_ops] rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${TAG}_tests_ops.log | tail -30 ;;
    bnbench)
      timeout 300 python tools/bn_bench.py --stream 1 > gpurun_out/${TAG}_bnbench.log 2>&1
      timeout 300 python tools/bn_bench.py --stream 0 >> gpurun_out/${TAG}_bnbench.log 2>&1
      echo "[bnbench] rc=$?"; cat gpurun_out/${TAG}_bnbench.log ;;
    ncu_bn)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:bn_stream -c 6 -f \
        -o gpurun_out/${TAG}_bn_stream python tools/bn_bench.py --stream 1 --once > gpurun_out/${TAG}_ncu_bn.log 2>&1
      echo "[ncu_bn] rc=$?" ;;
    attnbench)
      timeout 200 python tools/attn_bench.py > gpurun_out/${TAG}_attnbench.log 2>&1
      echo "[attnbench] rc=$?"; cat gpurun_out/${TAG}_attnbench.log
      timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_ -c 3 -f \
        -o gpurun_out/${TAG}_attn python tools/attn_bench.py --once > gpurun_out/${TAG}_ncu_attn.log 2>&1
      echo "[ncu_attn] rc=$?" ;;
    ddp2|ddp4|ddp8)
      N=${part#ddp}
      for cfg in "default" "CRIS_B200_BWD_SEGMENTS=3" "CRIS_B200_FUSED_SYNCBN=0"; do
        envs=""; [ "$cfg" != "default" ] && envs="$cfg"
        env $envs timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
          bench.py --gpus $N --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-incumbent > gpurun_out/${TAG}_ddp${N}_${cfg%%=*}.json 2> gpurun_out/${TAG}_ddp${N}.err
        echo "[ddp$N $cfg] rc=$?"; python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/${TAG}_ddp${N}_${cfg%%=*}.json').read().strip().splitlines()[-1]); print('   ', d['value'], 'img/s', d['ms_per_step'], 'ms/step  e2e', d['e2e']['value'])
except Exception as e: print('   parse error', e)
"
      done; tail -3 gpurun_out/${TAG}_ddp${N}.err ;;
    ddpq2|ddpq4|ddpq8)
      N=${part#ddpq}
      timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 \
        bench.py --gpus $N --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-incumbent > gpurun_out/${TAG}_ddpq${N}.json 2> gpurun_out/${TAG}_ddpq${N}.err
      echo "[ddpq$N] rc=$?"; tail -c 1500 gpurun_out/${TAG}_ddpq${N}.json; tail -3 gpurun_out/${TAG}_ddpq${N}.err ;;
    ddpbreak2|ddpbreak4|ddpbreak8)
      N=${part#ddpbreak}
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
        tools/ddp_breakdown.py > gpurun_out/${TAG}_ddpbreak${N}.log 2>&1
      echo "[ddpbreak$N] rc=$?"; grep -E "fwd|FAILED" gpurun_out/${TAG}_ddpbreak${N}.log ;;
    bwdover)
      timeout 300 python tools/bwd_overhead.py > gpurun_out/${TAG}_bwdover.log 2>&1
      echo "[bwdover] rc=$?"; tail -3 gpurun_out/${TAG}_bwdover.log ;;
    sweep)
      timeout 900 python tools/flag_sweep.py 64 > gpurun_out/${TAG}_sweep.log 2>&1
      echo "[sweep] rc=$?"; cat gpurun_out/${TAG}_sweep.log | tail -12 ;;
    tests_attn)
      timeout 150 python -m pytest tests/test_ops_gpu.py -k "attention" -q --no-header -p no:cacheprovider -rA -x \
        > gpurun_out/${TAG}_tests_attn.log 2>&1
      echo "[tests_attn] rc=$?"; grep -E "^(PASSED|FAILED|ERROR)|passed|failed|assert|Error|timeout|cris_b200" gpurun_out/${TAG}_tests_attn.log | tail -30 ;;
    tests_peer)
      timeout 600 python -m pytest tests/test_peer_gpu.py -q --no-header -p no:cacheprovider -rA > gpurun_out/${TAG}_tests_peer.log 2>&1
      echo "[tests_peer] rc=$?"; grep -E "^(PASSED|FAILED|ERROR)|passed|failed|assert|Error" gpurun_out/${TAG}_tests_peer.log | tail -10 ;;
    tests_post)
      timeout 300 python -m pytest tests/test_postproc_gpu.py tests/test_ops_gpu.py -k "postproc or bn_stream or conv_bn" -q --no-header -p no:cacheprovider -rA \
        > gpurun_out/${TAG}_tests_post.log 2>&1
      echo "[tests_post] rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|assert" gpurun_out/${TAG}_tests_post.log | tail -20 ;;
    tests_feeder)
      timeout 400 python -m pytest tests/test_feeder_gpu.py tests/test_ops_gpu.py -k "feeder or letterbox or realistic" -q --no-header -p no:cacheprovider -rA \
        > gpurun_out/${TAG}_tests_feeder.log 2>&1
      echo "[tests_feeder] rc=$?"; grep -E "^(PASSED|FAILED|ERROR)|passed|failed|assert|Error" gpurun_out/${TAG}_tests_feeder.log | tail -30 ;;
    wgstream)
      timeout 300 python -m pytest tests/test_parity_gpu.py -k "weight_gradient_branch or segmented" -q --no-header -p no:cacheprovider -rA \
        > gpurun_out/${TAG}_tests_wgstream.log 2>&1
      echo "[tests_wgstream] rc=$?"; grep -E "^(PASSED|FAILED|ERROR)|passed|failed|assert|Error" gpurun_out/${TAG}_tests_wgstream.log | tail -12
      CRIS_SWEEP="CRIS_B200_WGRAD_STREAM=1+CRIS_B200_WGRAD_JOIN=1000,CRIS_B200_WGRAD_STREAM=0,CRIS_B200_WGRAD_STREAM=1+CRIS_B200_WGRAD_JOIN=64,CRIS_B200_WGRAD_STREAM=0+CRIS_X=1,CRIS_B200_WGRAD_STREAM=1+CRIS_B200_WGRAD_JOIN=1000+CRIS_X=1" \
        timeout 600 python tools/flag_sweep.py 64 > gpurun_out/${TAG}_wgstream_sweep.log 2>&1
      echo "[wgstream sweep] rc=$?"; cat gpurun_out/${TAG}_wgstream_sweep.log | tail -8 ;;
    wg3)
      timeout 300 python -m pytest tests/test_gemm_native_gpu.py tests/test_ops_gpu.py -k "native or conv_bn or halo or deterministic" -q --no-header -p no:cacheprovider -rA \
        > gpurun_out/${TAG}_tests_wg3.log 2>&1
      echo "[tests_wg3] rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|assert|Error|MISMATCH|FAIL" gpurun_out/${TAG}_tests_wg3.log | tail -12
      for i in 13 16 17; do
        for f in 1 0; do echo "WGRAD3=$f"; CRIS_B200_WGRAD3=$f timeout 100 tests/native/gemm_selftest perf $i 2>&1 | grep PERF; done
      done ;;
    tests_optim)
      timeout 300 python -m pytest tests/test_optim_gpu.py tests/test_syncbn_equiv_gpu.py -q --no-header -p no:cacheprovider -rA > gpurun_out/${TAG}_tests_optim.log 2>&1
      echo "[tests_optim] rc=$?"; grep -E "^(PASSED|FAILED|ERROR)|passed|failed" gpurun_out/${TAG}_tests_optim.log | tail -8 ;;
    tests_new)
      timeout 1200 python -m pytest tests/test_parity_full_gpu.py tests/test_syncbn_equiv_gpu.py tests/test_dropin_gpu.py -q --no-header \
        -p no:cacheprovider -rA -s > gpurun_out/${TAG}_tests_new.log 2>&1
      echo "[tests_new] rc=$?"; grep -E "^(PASSED|FAILED|ERROR|SKIPPED)|passed|failed|norm_ratio|^loss|gradient tensors" gpurun_out/${TAG}_tests_new.log | tail -80 ;;
    bench)
      timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
      echo "[bench] rc=$?"; tail -c 6000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err ;;
    benchq)
      timeout 600 python bench.py --steps 10 --warmup 3 --no-incumbent --no-cpu-baseline > gpurun_out/${TAG}_benchq.json 2> gpurun_out/${TAG}_benchq.err
      echo "[benchq] rc=$?"; tail -c 5000 gpurun_out/${TAG}_benchq.json; tail -5 gpurun_out/${TAG}_benchq.err ;;
    ncu_gemm)
      # captures stay on the box (each --set full report is ~10 MB, gpurun_out may bring back 64 MiB): only the summary returns
      reps=""
      for i in 3 6 11 13 14 15; do
        timeout 200 ncu --set full --clock-control none -k regex:gemm_tc -c 1 -f \
          -o /tmp/${TAG}_gemm_perf$i tests/native/gemm_selftest perf $i > gpurun_out/${TAG}_ncu_gemm$i.log 2>&1
        echo "[ncu gemm perf $i] rc=$?"; reps="$reps /tmp/${TAG}_gemm_perf$i.ncu-rep"
      done
      timeout 200 ncu --set full --clock-control none -k regex:"conv_halo|layernorm_bwd|adam|dynconv_bce_fwd|upsample2x_fwd" -c 6 -f \
        -o /tmp/${TAG}_misc python tools/profile_step.py 64 > gpurun_out/${TAG}_ncu_misc.log 2>&1
      echo "[ncu misc] rc=$?"; reps="$reps /tmp/${TAG}_misc.ncu-rep"
      python tools/ncu_summary.py $reps > gpurun_out/${TAG}_ncu_gemm_summary.md 2>&1
      cp /tmp/${TAG}_gemm_perf3.ncu-rep gpurun_out/ 2>/dev/null
      cat gpurun_out/${TAG}_ncu_gemm_summary.md ;;
    ncu_wgrad)
      for i in 11 12 14; do
        timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 1 -f \
          -o gpurun_out/${TAG}_wgrad_perf$i tests/native/gemm_selftest perf $i > gpurun_out/${TAG}_ncu_wgrad$i.log 2>&1
        echo "[ncu wgrad $i] rc=$?"
      done ;;
    launches)
      timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py 64 > gpurun_out/${TAG}_launches.log 2>&1
      echo "[launches] rc=$?"; python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_summary.txt 2>&1
      head -45 gpurun_out/${TAG}_launches_summary.txt ;;
  esac
done
