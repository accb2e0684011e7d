#!/bin/bash
# Timing experiment: where does a tile of k_mlp_bwd_tc2 spend its time?  The four heads' backward of one real training step is
# re-timed with pieces of the kernel switched off at run time (al_set_bwd_debug; results are wrong while a bit is set, the
# model is trained and the step is built with the normal kernel):  bit 0 = no epilogues, bit 1 = no GEMMs, bit 2 = no
# output-gradient assembly / d-x write-out.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AL_BWD_DBG_SWEEP=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --c5-steps 0 --no-early-leg --render-frames 0 \
    > gpurun_out/bench_bwd_dbg.json 2> gpurun_out/bench_bwd_dbg.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_bwd_dbg.json'):
    if l.startswith('{'):
        d = json.loads(l); p = d['phases_ms']
        print('samples', p['live_samples'], 'field_backward', round(p['field_backward'], 4))
        for k, v in p['mlp_backward_debug_sweep_ms'].items():
            print(f"  bits {k} (epilogues {'off' if int(k) & 1 else 'on '}, GEMMs {'off' if int(k) & 2 else 'on '}, assembly+dx {'off' if int(k) & 4 else 'on '}): {v:.4f} ms")
PY
tail -2 gpurun_out/bench_bwd_dbg.err
