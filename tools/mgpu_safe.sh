#!/bin/bash
# tools/mgpu_safe.sh <ngpus> <size> : refuse to start when the host cannot hold one global mesh per rank
N=$1; n=$2
need=$(( N * (n*n*n*6/1000000) * 300 / 1000 + 20 ))   # ~0.3 GB per million cells per rank + slack, in GB
avail=$(free -g | awk '/Mem:/{print $7}')
echo "host: $(nproc) cpus, ${avail} GB available, need ~${need} GB"
if [ "$avail" -lt "$need" ]; then echo "not enough host memory, skipping"; exit 0; fi
tools/mgpu.sh $N $n
