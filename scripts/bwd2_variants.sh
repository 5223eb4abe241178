#!/bin/bash
# tuning: the CTA-pair attention backward at the bench shape under the host-side switches
for v in "" "VDS_BWD2_TMA_OUT=0" "VDS_BWD2_UNIFORM=3" "VDS_BWD2_UNIFORM=3 VDS_BWD2_TMA_OUT=0" "VDS_BWD2_UNIFORM=2" "VDS_BWD2_UNIFORM=4"; do
  echo "== $v"
  env $v python scripts/attn_bwd_pair_check.py 1 2>&1 | grep "pair_mode=1"
done
