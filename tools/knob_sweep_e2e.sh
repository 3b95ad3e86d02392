#!/bin/bash
# host-asynchronous path: views per sub-chunk x auxiliary streams
cd "$(dirname "$0")/.."
for c in 2 4 8 16; do for a in 1 2 3; do
  echo -n "async_chunk $c aux_host $a  "
  B2R_ASYNC_CHUNK=$c B2R_AUX_HOST=$a python tools/e2e_async_probe.py 2>&1 | tail -1
done; done
