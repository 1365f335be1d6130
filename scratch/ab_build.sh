#!/bin/bash
# A/B support: export HEAD into scratch/ab_base (git-ignored, travels with gpurun) and build its library there,
# so that one gpurun call can time the committed code against the working tree on the SAME box.
set -e
cd "$(dirname "$0")/.."
rm -rf scratch/ab_base && mkdir -p scratch/ab_base
git archive HEAD | tar -x -C scratch/ab_base
rm -rf scratch/ab_base/scratch scratch/ab_base/profiles
(cd scratch/ab_base && python -c "
from arreau_b200 import build as b
print(b.build(force=True))")
