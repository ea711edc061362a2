#!/bin/bash
for s in 1 2; do echo "TO_EXPAND_SEEDS=$s"; TO_EXPAND_SEEDS=$s python profiles/bench_phases.py --steps 10 --warmup 3 --no-e2e; done
