# Records shared by two initial sectors (QIW_LANE_DUAL): the previous build, the new build with the
# feature off (code-generation effect alone) and on; README run, then the step kernel with the machine full.
for rep in 1 2; do
  echo "== previous build (pass $rep)";            QIW_LIB=$PWD/variants/v1.so   timeout 300 python profiles/run_vs_step.py 200 1024 2>&1 | head -1
  echo "== new build, QIW_LANE_DUAL=0 (pass $rep)"; QIW_LANE_DUAL=0 timeout 300 python profiles/run_vs_step.py 200 1024 2>&1 | head -1
  echo "== new build (pass $rep)";                  timeout 300 python profiles/run_vs_step.py 200 1024 2>&1 | head -1
done
echo "== orders 0:4, previous / new"
QIW_LIB=$PWD/variants/v1.so timeout 300 python profiles/throughput.py 4 131072 2>&1 | tail -1
timeout 300 python profiles/throughput.py 4 131072 2>&1 | tail -1
echo "== orders 0:6, previous / new"
QIW_LIB=$PWD/variants/v1.so timeout 300 python profiles/throughput.py 6 16384 2>&1 | tail -1
timeout 300 python profiles/throughput.py 6 16384 2>&1 | tail -1
