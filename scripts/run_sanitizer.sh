#!/bin/bash
# compute-sanitizer over scripts/sanitizer_smoke.py (every kernel family incl. the three PT launch shapes).
# racecheck does not understand flag-based hand-over (the producer/consumer ring of the warp-specialised PT kernel
# synchronises through volatile step counters + __threadfence_block, not barriers), so that kernel is left out of
# the racecheck pass and covered by memcheck/synccheck and by the bit-identity tests.
out=${1:-gpurun_out/sanitizer.txt}
: > $out
for tool in memcheck synccheck; do
    echo "\$ compute-sanitizer --tool $tool --error-exitcode 7 python scripts/sanitizer_smoke.py" >> $out
    timeout 1500 compute-sanitizer --tool $tool --error-exitcode 7 python scripts/sanitizer_smoke.py 2>&1 | grep -E "=========|^ok|Error|error" | tail -12 >> $out
    echo "(exit code ${PIPESTATUS[0]})" >> $out
done
echo "\$ SANITIZER_SKIP_HELP=1 CARMA_PT_HELP=0 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/sanitizer_smoke.py" >> $out
SANITIZER_SKIP_HELP=1 CARMA_PT_HELP=0 timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/sanitizer_smoke.py 2>&1 | grep -E "=========|^ok|Error|error" | tail -12 >> $out
echo "(exit code ${PIPESTATUS[0]})" >> $out
cat $out
