#!/bin/bash
# latency (small batch) vs throughput (large batch) of every phase kernel
for b in 148 592 1184 2368 4096 8192; do echo "batch $b"; python profiles/bench_phases.py --steps 10 --warmup 3 --no-e2e --batch $b; done
