#!/bin/bash
# rebuild csrc/libmmlrec_b200.so in-tree (from any working directory); fails loudly
set -e
cd "$(dirname "$0")/.."
python -c "from mmlrec_b200.csrc.build import build; print(build(force=False, verbose=False))"
