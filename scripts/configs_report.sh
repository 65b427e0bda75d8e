#!/bin/bash
# Every BASELINE.json config beyond the headline bench line, measured in one GPU-box visit.
# Output: gpurun_out/configs.log (copied to profiles/ per round).
set -u
out=gpurun_out/configs.log
mkdir -p gpurun_out
{
echo "## config 1-3 + pla85900: NN(0) tour -> 2-opt local optimum (scripts/pruned_time.py)"
python scripts/pruned_time.py berlin52 pr2392 usa13509 d15112 pla85900 2>&1 | cut -c1-260
echo
echo "## configs 2 and 4: NN and multi-start NN + 2-opt (scripts/nn_time.py)"
python scripts/nn_time.py 2>&1 | cut -c1-260
echo
echo "## config 5: tabu / VNS on d15112 and the 20 000-node random instance (scripts/config5_time.py 1000)"
python scripts/config5_time.py 1000 2>&1 | cut -c1-400
echo
echo "## K1 cost matrix build (scripts/matrix_time.py)"
python scripts/matrix_time.py 2>&1 | cut -c1-200
echo
echo "## Extra Mileage (scripts/em_time.py)"
python scripts/em_time.py rl1304 pr2392 usa13509 d15112 2>&1 | cut -c1-200
} > $out 2>&1
cat $out
