#!/bin/bash
# sweep the Riccati occupancy / ring-depth variants (TO_RICCATI_VARIANT, see riccati.cu)
for v in 0 1 2 3 4; do TO_RICCATI_VARIANT=$v python profiles/bench_phases.py --steps 10 --warmup 3 --no-e2e; done
