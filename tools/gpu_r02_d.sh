#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity.log
( timeout 900 python tests/tools/diag_engine.py 512 1 ) > $O/r02d_diag_512_b1.log 2>&1
( timeout 600 python tests/tools/diag_engine.py 256 1 ) > $O/r02d_diag_256_b1.log 2>&1
( timeout 600 python -m pytest tests -m gpu -q -k "instance_norm or to_one or graph" ) > $O/r02d_pytest.log 2>&1
cp $O/parity.log $O/r02d_parity.log 2>/dev/null
tail -12 $O/r02d_diag_512_b1.log; tail -12 $O/r02d_diag_256_b1.log; tail -12 $O/r02d_pytest.log; grep instance_norm $O/r02d_parity.log
