#!/bin/bash
# host ingest alone (no GPU work): reader threads -> parsed batches, on a synthetic config[1] FASTQ in tmpfs
set -e
g++ -O2 -std=c++17 -pthread tests/cpp/ingest_check.cpp -lz -o /tmp/ingest_check
python profiles/f2f_bench.py 20000 .fq gzin prepare_only > /dev/null 2>&1 || true
cp /dev/shm/f2f/in.fq.gz /dev/shm/f2f/in16.fq.gz
python profiles/f2f_bench.py 20000 .fq bgzfin prepare_only > /dev/null 2>&1 || true
ls -la /dev/shm/f2f/
TIMEFORMAT="%R s"
t() { time "$@" > /dev/null; }
echo -n "plain, parallel chunk parser (8 thr): "; INGEST_ONLY=parallel t /tmp/ingest_check /dev/shm/f2f/in.fq 1 134217728 8
echo -n "plain, serial reader:                 "; INGEST_ONLY=serial t /tmp/ingest_check /dev/shm/f2f/in.fq 1 67108864 1
echo -n "gzip 16 members, own inflate, 1 thr:  "; TGSF_INFLATE_THREADS=1 INGEST_ONLY=serial t /tmp/ingest_check /dev/shm/f2f/in16.fq.gz 1 67108864 1
echo -n "gzip 16 members, zlib:                "; TGSF_ZLIB_INFLATE=1 INGEST_ONLY=serial t /tmp/ingest_check /dev/shm/f2f/in16.fq.gz 1 67108864 1
for n in 1 2 4 8; do echo -n "gzip 16 members, $n inflate threads:   "; TGSF_INFLATE_THREADS=$n INGEST_ONLY=serial t /tmp/ingest_check /dev/shm/f2f/in16.fq.gz 1 67108864 1; done
for n in 1 2 4 8; do echo -n "BGZF, $n inflate threads:              "; TGSF_INFLATE_THREADS=$n INGEST_ONLY=serial t /tmp/ingest_check /dev/shm/f2f/in.fq.gz 1 67108864 1; done
# one deflate stream (plain `gzip`): two-pass parallel decode, src/pinflate.hpp
zcat /dev/shm/f2f/in16.fq.gz | gzip -1 > /dev/shm/f2f/in1.fq.gz
for n in 1 2 4 8; do echo -n "gzip single member, $n inflate threads: "; TGSF_INFLATE_THREADS=$n INGEST_ONLY=serial t /tmp/ingest_check /dev/shm/f2f/in1.fq.gz 1 67108864 1; done
