#!/bin/bash
# Runs every UMMA probe case in its own process (see umma_probe.cu); log -> gpurun_out/umma_probe.log
cd "$(dirname "$0")"
mkdir -p ../gpurun_out
LOG=../gpurun_out/umma_probe.log
: > $LOG
while read -r M N K A B BULK; do
  [ -z "$M" ] && continue
  timeout 30 ./umma_probe $M $N $K $A $B $BULK >> $LOG 2>&1 || echo "  -> exit $? for $M $N $K $A $B $BULK" >> $LOG
done <<CASES
128 128 64 0 0 0
128 128 128 0 0 0
128 48 128 0 0 0
128 16 128 0 0 0
128 64 16 0 0 0
128 64 48 0 0 0
128 64 80 0 0 0
128 128 16 0 1 0
128 128 128 0 1 0
128 48 128 0 1 0
128 16 64 0 1 0
128 128 128 1 1 0
128 48 128 1 1 0
128 16 128 1 1 0
64 64 128 1 1 0
64 16 128 1 1 0
64 80 128 1 1 0
64 64 64 0 0 0
128 128 64 0 0 1
128 128 128 1 1 1
CASES
cat $LOG
