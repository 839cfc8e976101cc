#!/bin/bash
# AddressSanitizer + UBSan build of the C++ host against the CUDA library, a few representative runs
set -e
cd "$(dirname "$0")/.."
g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-omit-frame-pointer -o /tmp/tgsfilter_asan src/TGSFilter.cpp \
    -Ltgsfilter_b200 -ltgsf_cuda -lz -pthread -Wl,-rpath,$PWD/tgsfilter_b200
python - <<'PY'
import os, sys, gzip
sys.path.insert(0, os.getcwd())
from tgsfilter_b200 import synth
os.makedirs("/dev/shm/asan", exist_ok=True)
b = synth.make_config(2, 1500, max_len=60000)
fq = b.to_fastq()
open("/dev/shm/asan/in.fq", "wb").write(fq)
open("/dev/shm/asan/in.fq.gz", "wb").write(gzip.compress(fq[:len(fq)//2], 1) + gzip.compress(fq[len(fq)//2:], 6))
b5 = synth.make_config(5, 800, max_len=20000)
open("/dev/shm/asan/hifi.fq", "wb").write(b5.to_fastq())
PY
export ASAN_OPTIONS=protect_shadow_gap=0:detect_leaks=0 TGSF_CLEAN_EXIT=1 TGSF_BATCH_MB=4
run() { echo "== $*"; /tmp/tgsfilter_asan "$@" 2>&1 | grep -E "ERROR|runtime error|SUMMARY|Sanitizer" | head -5 || true; }
run -i /dev/shm/asan/in.fq -x ont -o /dev/shm/asan/o1.fq
run -i /dev/shm/asan/in.fq -x ont -o /dev/shm/asan/o2.fq.gz
run -i /dev/shm/asan/in.fq.gz -x ont -o /dev/shm/asan/o3.fq.gz
run -i /dev/shm/asan/hifi.fq -x hifi -k 11 -p 40 -g 2m -d 3 -o /dev/shm/asan/o4.fq
run -i /dev/shm/asan/hifi.fq -x hifi -f -o /dev/shm/asan/o5.fa
run -i /dev/shm/asan/hifi.fq --qc
run -i /dev/shm/asan/hifi.fq -F -r 100 -o /dev/shm/asan/o7.fq.gz
TGSF_PREPASS_BUFFER_MB=2 run -i /dev/shm/asan/in.fq.gz -x ont -o /dev/shm/asan/o6.fq
ls -la /dev/shm/asan | head -12
