#!/bin/bash
# k_bin grid / quad-sharing sweep: stage times of the 16-view profile step
for blocks in 1184 592 296 148 74 37; do for share in 1 4 16 64; do
  echo -n "blocks=$blocks share=$share  "
  B2R_BIN_BLOCKS=$blocks B2R_BIN_SHARE=$share python tools/profile_step.py 16 4 2>&1 | grep "^3 " 
done; done
