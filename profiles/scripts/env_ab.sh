# A/B of run-kernel launch parameters (environment switches of the library) on the README run.
for rep in 1 2; do
for pw in 4 5 6 7 8; do
  echo "== QIW_RUN_POST_WARPS=$pw (pass $rep)"
  QIW_RUN_POST_WARPS=$pw timeout 300 python profiles/run_vs_step.py 200 1024 2>&1 | head -1
done
done
