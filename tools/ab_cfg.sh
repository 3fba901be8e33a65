for cfgline in "2048 152 2 20 14 2 2048" "2048 152 2 2 14 4 2048" "1024 72 4 2 14 6 1024 mmse" "4096 288 1 2 14 4 2048" "512 36 4 2 14 4 4096"; do
  for L in lib_old.so librubmimo_b200.so; do echo -n "$L: "; RUB_MIMO_LIB=$PWD/rub_mimo_b200/$L timeout 100 python tools/time_config.py $cfgline 2>&1 | tail -1; done
done
