# launch-shape sweep of the two GEMVs on the C2 window (timing mode, CUDA events per kernel)
python tests/ab_probe.py gemvVariantF 0 1 2 3 4 5 6 7 8 2>&1 | grep -v "^$" | cut -c1-260
python tests/ab_probe.py gemvVariantB 0 1 2 3 4 5 6 7 2>&1 | cut -c1-260
python tests/ab_probe.py gemvGridMul 4 8 12 16 2>&1 | cut -c1-260
