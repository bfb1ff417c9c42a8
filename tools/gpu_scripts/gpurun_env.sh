# stage times under environment switches: ENVS="A=1 B=2;C=3" (';' separates runs), SCENES="sphere bench"
IFS=';' read -ra RUNS <<< "${ENVS:-}"
for sc in ${SCENES:-sphere}; do python tools/stage_probe.py $sc 0 | grep flags; done
for r in "${RUNS[@]}"; do echo "== $r"; for sc in ${SCENES:-sphere}; do env $r python tools/stage_probe.py $sc 0 | grep flags; done; done
