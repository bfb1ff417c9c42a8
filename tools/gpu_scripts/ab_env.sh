# A/B of one library under environment switches + the GPU tests: ENVS="A=1;B=1" SCENES="sphere bench"
if [ "${TESTS:-1}" = 1 ]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -3; fi
IFS=';' read -ra RUNS <<< "${ENVS:-}"
for rep in 1 2; do
echo "== default"; for sc in ${SCENES:-sphere}; do python tools/stage_probe.py $sc 0 | grep flags; done
for r in "${RUNS[@]}"; do echo "== $r"; for sc in ${SCENES:-sphere}; do env $r python tools/stage_probe.py $sc 0 | grep flags; done; done
done
