# A/B of library builds (variants/*.so, made by hand from modified copies of csrc/) on the README run: device time of
# the run kernel.  usage: bash profiles/scripts/variants_1gpu.sh v0 v1 ...
for rep in 1 2; do
for v in "$@"; do
  echo "== $v (pass $rep)"
  QIW_LIB=$PWD/variants/$v.so timeout 300 python profiles/run_vs_step.py 200 1024 2>&1 | head -1
done
done
