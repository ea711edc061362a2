#!/bin/bash
# sweep the Riccati occupancy / ring-depth variants (TO_RICCATI_VARIANT, see riccati.cu)
for v in 1 3 5 6; do TO_RICCATI_VARIANT=$v python profiles/bench_phases.py --steps 10 --warmup 3 --no-e2e; done
