#!/bin/bash
# host-asynchronous path: views per sub-chunk x auxiliary streams (64-view batches)
cd "$(dirname "$0")/.."
for c in 4 8 16 32 64; do for a in 1 2; do
  echo -n "async_chunk $c aux_host $a  "
  B2R_ASYNC_CHUNK=$c B2R_AUX_HOST=$a python tools/e2e_async_probe.py 64 2>&1 | tail -1
done; done
