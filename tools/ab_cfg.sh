#!/bin/bash
# A/B of two builds over a list of configurations (tools/time_config.py arguments), e.g.
#   CFGS="4096 288 1 2 14 4 2048;2048 152 2 2 14 4 2048" tools/ab_cfg.sh lib_old.so librubmimo_b200.so
IFS=';' read -ra LIST <<< "${CFGS:-4096 288 1 2 14 4 2048;4096 288 1 2 14 6 2048;2048 152 2 2 14 4 2048}"
for cfgline in "${LIST[@]}"; do
  for L in "$@"; do echo -n "$L: "; RUB_MIMO_LIB=$PWD/rub_mimo_b200/$L timeout 100 python tools/time_config.py $cfgline 2>&1 | tail -1; done
done
