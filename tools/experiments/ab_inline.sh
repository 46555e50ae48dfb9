# A/B of experiment builds of the pairing kernels (libtcb200_<tag>.so next to the default library): bash tools/experiments/ab_inline.sh "_v3 _v4"
for v in "" $1; do
  L=threshold_crypto_b200/csrc/libtcb200$v.so
  TCB200_LIB=$L python tools/kbench.py r2w$v verify 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['lib'], 'hash', round(d['hash_g2_2^16']['best_ms'],2), 'pairing', round(d['pairing_2^16']['best_ms'],2), 'verify', round(d['verify_2^16']['best_ms'],2))"
done
