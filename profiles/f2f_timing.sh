#!/bin/bash
# phase timings of the C++ host on a synthetic config[1] FASTQ (tmpfs)
python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from tgsfilter_b200 import synth
b = synth.make_config(2, 20000, with_names=False)
os.makedirs("/dev/shm/f2f", exist_ok=True)
open("/dev/shm/f2f/in.fq", "wb").write(b.to_fastq())
PY
for i in 1 2; do TGSF_TIMING=1 ./src/tgsfilter -i /dev/shm/f2f/in.fq -x ont -o /dev/shm/f2f/o.fq 2>&1 | grep timing; echo; done
TGSF_BATCH_MB=32 TGSF_TIMING=1 ./src/tgsfilter -i /dev/shm/f2f/in.fq -x ont -o /dev/shm/f2f/o.fq 2>&1 | grep timing
