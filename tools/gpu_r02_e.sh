#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python tests/tools/diag_engine.py 256 1 ) > $O/r02e_diag_256_b1.log 2>&1
( timeout 600 python tests/tools/diag_engine.py 256 2 ) > $O/r02e_diag_256_b2.log 2>&1
( timeout 600 python tests/tools/diag_engine.py 128 1 ) > $O/r02e_diag_128_b1.log 2>&1
for f in $O/r02e_diag_256_b1.log $O/r02e_diag_256_b2.log $O/r02e_diag_128_b1.log; do echo == $f; grep -E "dual_up3|model.8|head.dx|last.dx|kernel dy|second kernel" $f; done
