# A/B of experiment builds of the G1 kernels: bash tools/experiments/ab_g1.sh "_m4"
for v in "" $1; do
  L=threshold_crypto_b200/csrc/libtcb200$v.so
  TCB200_LIB=$L python tools/kbench.py r2g1$v decrypt 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['lib'], {k:round(v['best_ms'],2) for k,v in d.items() if isinstance(v,dict)})"
done
