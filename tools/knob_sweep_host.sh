#!/bin/bash
# synchronous host output (render_batch(out=pinned)): views per sub-chunk x auxiliary streams
cd "$(dirname "$0")/.."
for c in 1 2 4 8; do for a in 1 2 3; do
  echo -n "host_chunk $c aux_host $a  "
  B2R_HOST_CHUNK=$c B2R_AUX_HOST=$a python tools/e2e_breakdown.py 2>&1 | grep "render_packed -> host"
done; done
