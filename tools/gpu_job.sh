#!/bin/bash
# One parameterised GPU job (replaces the per-round scripts):   gpurun -- 'bash tools/gpu_job.sh TAG step [step ...]'
# TAG names the outputs under gpurun_out/ (e.g. r2a).  Steps:
#   tests [pytest args]   GPU parity tests (-m gpu), per-test timeout            -> pytest_TAG.log
#   smoke                 __graft_entry__.smoke()                                -> smoke_TAG.log
#   bench [bench args]    python bench.py ...                                    -> bench_TAG.json / .err
#   reference [args]      python bench.py --impl reference ...                   -> bench_TAG_reference.json
#   launches              ncu launch list of two training steps                  -> launches_TAG.csv / .md
#   launches_render       ncu launch list of one full-frame render                -> launches_render_TAG.csv / .md
#   launches_c5           ncu launch list of two training steps of the C5 configuration -> launches_c5_TAG.csv / .md
#   ncufull [regex]       ncu --set full --import-source of one step (no graph)  -> step_TAG.ncu-rep, ncu_full_TAG.csv, traffic
#   sanitize              compute-sanitizer memcheck + racecheck + synccheck on tools/sanitize_target.py -> sanitize_TAG_{tool}.log
#   probe                 tools/tmem_probe (TMEM read throughput)                -> tmem_probe_TAG.log
#   exchange              tools/bench_exchange.py under torchrun (needs --gpus N) -> exchange_TAG.json
# A step's own arguments follow it up to the next step name.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=$1; shift
STEPS="tests smoke bench reference launches launches_render launches_c5 ncufull sanitize probe exchange"
is_step() { for s in $STEPS; do [ "$1" = "$s" ] && return 0; done; return 1; }
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_$TAG.txt
while [ $# -gt 0 ]; do
  step=$1; shift
  args=()
  while [ $# -gt 0 ] && ! is_step "$1"; do args+=("$1"); shift; done
  echo "== $step ${args[*]}"
  case $step in
    tests)
      rm -f gpurun_out/parity_report.json
      timeout -k 5 ${AL_TEST_TIMEOUT:-420} python -m pytest tests -m gpu -q --timeout 120 "${args[@]}" > gpurun_out/pytest_$TAG.log 2>&1
      echo "pytest exit $?" >> gpurun_out/pytest_$TAG.log
      grep -E "^(FAILED|ERROR)|passed|failed|pytest exit" gpurun_out/pytest_$TAG.log | tail -12
      [ -f gpurun_out/parity_report.json ] && cp gpurun_out/parity_report.json gpurun_out/parity_report_$TAG.json ;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log ;;
    bench)
      timeout 1500 python bench.py "${args[@]}" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
      echo "bench exit $?"; tail -3 gpurun_out/bench_$TAG.err
      python tools/show_bench.py gpurun_out/bench_$TAG.json ;;
    reference)
      timeout 1500 python bench.py --impl reference "${args[@]}" > gpurun_out/bench_${TAG}_reference.json 2>> gpurun_out/bench_$TAG.err
      cut -c1-400 gpurun_out/bench_${TAG}_reference.json ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
          --log-file gpurun_out/launches_$TAG.csv python bench.py --ncu-range 2 --no-cpu-baseline "${args[@]}" > gpurun_out/launch_$TAG.log 2>&1
      python tools/summarize_ncu.py launches gpurun_out/launches_$TAG.csv > gpurun_out/launches_$TAG.md 2>&1; head -40 gpurun_out/launches_$TAG.md ;;
    launches_render)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
          --log-file gpurun_out/launches_render_$TAG.csv python bench.py --ncu-render 1 --no-cpu-baseline "${args[@]}" > gpurun_out/launch_render_$TAG.log 2>&1
      python tools/summarize_ncu.py launches gpurun_out/launches_render_$TAG.csv > gpurun_out/launches_render_$TAG.md 2>&1; head -40 gpurun_out/launches_render_$TAG.md ;;
    launches_c5)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
          --log-file gpurun_out/launches_c5_$TAG.csv python bench.py --ncu-c5 2 --c5-pretrain 200 --no-cpu-baseline "${args[@]}" > gpurun_out/launch_c5_$TAG.log 2>&1
      python tools/summarize_ncu.py launches gpurun_out/launches_c5_$TAG.csv > gpurun_out/launches_c5_$TAG.md 2>&1; head -40 gpurun_out/launches_c5_$TAG.md ;;
    ncufull)
      K=(); [ ${#args[@]} -gt 0 ] && K=(-k "regex:${args[0]}")
      AL_NO_GRAPH=1 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off "${K[@]}" \
          -o gpurun_out/step_$TAG -f python bench.py --ncu-range 1 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
      tail -2 gpurun_out/ncu_full_$TAG.log; ls -la gpurun_out/step_$TAG.ncu-rep
      python tools/summarize_ncu.py full gpurun_out/step_$TAG.ncu-rep > gpurun_out/ncu_full_$TAG.csv 2> gpurun_out/ncu_full_$TAG.err
      python tools/summarize_ncu.py traffic gpurun_out/step_$TAG.ncu-rep > gpurun_out/roofline_traffic_$TAG.json 2>/dev/null
      head -3 gpurun_out/ncu_full_$TAG.csv
      # gpurun merges at most 64 MiB back: keep the summaries, drop a report that would push the whole directory over
      [ $(du -sm gpurun_out | cut -f1) -gt 55 ] && rm -f gpurun_out/step_$TAG.ncu-rep && echo "step_$TAG.ncu-rep dropped (size)" ;;
    sanitize)
      for tool in memcheck racecheck synccheck; do
        timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py \
            > gpurun_out/sanitize_${TAG}_$tool.log 2>&1
        echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|wide head ok|Error|hazard" gpurun_out/sanitize_${TAG}_$tool.log | sort | uniq -c | head -12
      done ;;
    probe)
      make -C tools tmem_probe > /dev/null 2>&1
      timeout 120 tools/tmem_probe > gpurun_out/tmem_probe_$TAG.log 2>&1; cat gpurun_out/tmem_probe_$TAG.log ;;
    exchange)
      N=$(nvidia-smi -L | wc -l)
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
          tools/bench_exchange.py "${args[@]}" > gpurun_out/exchange_$TAG.json 2> gpurun_out/exchange_$TAG.err
      cat gpurun_out/exchange_$TAG.json; tail -3 gpurun_out/exchange_$TAG.err ;;
    *) echo "unknown step $step" ;;
  esac
done
