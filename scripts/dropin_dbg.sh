#!/bin/bash
cd /root/repo
python - <<'PY'
import sys; sys.path.insert(0,'/root/repo')
from ahf_b200 import synth
b=synth.make_box(32, seed=21, n_clumps=6)
print(synth.write_reference_case(b,'/tmp/dd'))
PY
cd /tmp/dd
ulimit -c 0
if which gdb >/dev/null 2>&1; then
  gdb -batch -ex run -ex bt --args /root/repo/ahf_b200/host/_build/AHF-b200 AHF.input 2>&1 | tail -30
else
  AHFB200_VERBOSE=1 /root/repo/ahf_b200/host/_build/AHF-b200 AHF.input > out.log 2> err.log; echo rc=$?; tail -5 err.log; ls -la
fi
