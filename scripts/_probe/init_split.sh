#!/bin/bash
# start-up split of the drop-in program on a 128^3 box (context, module loads, first-touch costs), default and eager module loading
W=$(mktemp -d)
python - "$W" <<PY
import sys; sys.path.insert(0, ".")
from ahf_b200 import synth
box = synth.make_box(128, seed=43)
print(synth.write_reference_case(box, sys.argv[1]))
PY
cd $W
for mode in LAZY EAGER LAZY EAGER; do
  export CUDA_MODULE_LOADING=$mode
  t0=$(date +%s.%N)
  AHFGPU_INIT_TIMING=1 AHFB200_TIMING=1 $GRAFT_REPO_ROOT/ahf_b200/host/_build/AHF-b200-full AHF.input 2> err.txt >/dev/null
  t1=$(date +%s.%N)
  echo "$mode wall=$(echo "$t1 - $t0" | bc) rc=$?"; grep -a TIMING err.txt | sed "s/^/$mode /"
done
